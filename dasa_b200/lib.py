"""ctypes loader for libdasa_b200.so — the C-ABI library declared in include/dasa_b200.h.

The product path has no CPU or library fallback: if the shared object is missing or a call returns an error code,
this module raises. `check_exports()` is what the CPU test tier uses to verify that every symbol the header declares
is exported (no compute call is made without a GPU).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdasa_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "dasa_b200.h")

ERRORS = {-1: "BAD_SHAPE", -2: "BAD_ALIGN", -3: "WORKSPACE", -4: "CUDA", -5: "UNSUPPORTED"}

P, I, L, F, Z, U = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t, ctypes.c_uint64


class DasaError(RuntimeError):
    pass


class BiLstmFwd(ctypes.Structure):
    """dasa_bilstm_fwd_t"""
    _fields_ = [("xp", P * 2), ("w_hh", P * 2), ("b_ih", P * 2), ("b_hh", P * 2), ("hs", P * 2), ("cs", P * 2),
                ("acts", P * 2), ("out", P), ("lengths", P), ("B", I), ("L", I), ("H", I)]


class BiLstmBwd(ctypes.Structure):
    """dasa_bilstm_bwd_t"""
    _fields_ = [("w_hh_t", P * 2), ("acts", P * 2), ("cs", P * 2), ("dout", P), ("dh_fin", P * 2), ("dc_fin", P * 2),
                ("dgates", P * 2), ("dh_pass", P * 2), ("dc_work", P * 2), ("lengths", P), ("B", I), ("L", I), ("H", I)]


class BiLstmPackedFwd(ctypes.Structure):
    """dasa_bilstm_packed_fwd_t"""
    _fields_ = [("R", I), ("L", I), ("H", I), ("n_rows", P), ("off", P), ("perm", P), ("xp", P * 2), ("w_hh", P * 2), ("b_ih", P * 2),
                ("b_hh", P * 2), ("hprev", P * 2), ("cs", P * 2), ("acts", P * 2), ("out", P), ("h_fin", P * 2), ("c_fin", P * 2),
                ("out_mask", P), ("drop_seed_dev", P), ("drop_seed", U), ("drop_base", U), ("drop_p", F), ("drop_scale", F),
                ("w_hh16", P * 2), ("h16", P * 2)]


class BiLstmPackedBwd(ctypes.Structure):
    """dasa_bilstm_packed_bwd_t"""
    _fields_ = [("R", I), ("L", I), ("H", I), ("n_rows", P), ("off", P), ("perm", P), ("w_hh_t", P * 2), ("acts", P * 2), ("cs", P * 2),
                ("dout", P), ("dh_fin", P * 2), ("dc_fin", P * 2), ("dgates", P * 2), ("dc_work", P * 2),
                ("out_mask", P), ("drop_seed_dev", P), ("drop_seed", U), ("drop_base", U), ("drop_p", F), ("drop_scale", F),
                ("w_hh_t16", P * 2), ("dg16", P * 2)]


class DecoderFwd(ctypes.Structure):
    """dasa_decoder_fwd_t (field order and types exactly as in include/dasa_b200.h)"""
    _fields_ = ([(k, I) for k in ("T", "B", "H", "E", "F", "V", "L", "D", "headings", "shift_k", "NK")] +
                [("emb", P), ("feat", P), ("feat_ld_row", L), ("feat_ld_b", L), ("feat_ld_t", L),
                 ("ctx", P), ("ctx_ld_row", L), ("ctx_ld_b", L), ("ctx_ld_t", L), ("ctx_mask", P), ("ctx_mask_ld", L),
                 ("h0", P), ("c0", P), ("m_hprev", P), ("m_h1", P), ("drop_scale", F),
                 ("w_feat", P), ("b_feat", P), ("w_lstm", P), ("b_ih", P), ("b_hh", P), ("w_att_in", P), ("w_att_out", P)] +
                [(k, P) for k in ("hprev_drop", "tk", "p", "q", "kappa", "xh", "acts", "c", "h1", "cat", "t2", "alpha", "htilde",
                                  "zpart", "barrier", "x16")])


class DecoderBwd(ctypes.Structure):
    """dasa_decoder_bwd_t"""
    _fields_ = ([(k, I) for k in ("T", "B", "H", "E", "F", "V", "L", "D", "headings", "shift_k", "NK")] +
                [("feat", P), ("feat_ld_row", L), ("feat_ld_b", L), ("feat_ld_t", L),
                 ("ctx", P), ("ctx_ld_row", L), ("ctx_ld_b", L), ("ctx_ld_t", L), ("ctx_mask", P), ("ctx_mask_ld", L),
                 ("m_hprev", P), ("m_h1", P), ("drop_scale", F),
                 ("w_feat_t", P), ("ld_w_feat_t", L), ("w_lstm_t", P), ("ld_w_lstm_t", L), ("w_att_in_t", P), ("ld_w_att_in_t", L),
                 ("w_att_out_t", P), ("ld_w_att_out_t", L)] +
                [(k, P) for k in ("tk", "p", "q", "kappa", "acts", "c", "cat", "t2", "alpha", "htilde", "d_htilde", "d_h1",
                                  "d_c_last", "du", "dt2", "dgates", "dtk", "demb")] +
                [("dfeat", P), ("dfeat_ld_row", L), ("dfeat_ld_b", L), ("dfeat_ld_t", L)] +
                [(k, P) for k in ("dctx", "dh0", "dc0", "dcat", "dattn", "dhdir", "dc_carry", "zpart", "barrier", "g16")])


class Epilogue(ctypes.Structure):
    """dasa_epilogue_t"""
    _fields_ = [("bias", P), ("gate_src", P), ("ld_gate", L), ("gate_out", P), ("ld_gate_out", L),
                ("drop_mask", P), ("drop_scale", F)]


def _parse_header():
    """Derive every prototype (ctypes restype/argtypes) from include/dasa_b200.h, so the binding cannot drift from
    the declared ABI."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    sig = {}
    for m in re.finditer(r"(const\s+char\s*\*|size_t|int)\s+(dasa_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, params = m.group(1), m.group(2), m.group(3).strip()
        res = ctypes.c_char_p if "char" in ret else (Z if ret == "size_t" else I)
        args = []
        if params and params != "void":
            for prm in params.split(","):
                prm = prm.strip()
                if "dasa_epilogue_t" in prm:
                    args.append(ctypes.POINTER(Epilogue))
                elif "*" in prm:
                    args.append(P)
                elif re.match(r"(const\s+)?int64_t\b", prm):
                    args.append(L)
                elif re.match(r"(const\s+)?uint64_t\b", prm):
                    args.append(U)
                elif re.match(r"(const\s+)?size_t\b", prm):
                    args.append(Z)
                elif re.match(r"(const\s+)?float\b", prm):
                    args.append(F)
                elif re.match(r"(const\s+)?int\b", prm):
                    args.append(I)
                else:
                    raise DasaError("cannot map parameter %r of %s" % (prm, name))
        sig[name] = (res, args)
    return sig


_lib = None
launches = 0          # number of kernel-launching C-ABI calls made through `call` (bench.py reports it)


def header_symbols():
    """Names of every function declared in include/dasa_b200.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dasa_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DasaError("libdasa_b200.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                        "there is no fallback path" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _parse_header().items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check_exports():
    lib = load()
    names = header_symbols()
    missing = [s for s in names if not hasattr(lib, s)]
    unparsed = [s for s in names if s not in _parse_header()]
    return missing, unparsed


def call(name, *args):
    """Invoke a status-returning entry point; raise on a non-zero status."""
    global launches
    fn = getattr(load(), name)
    rc = fn(*args)
    launches += 1
    if rc != 0:
        raise DasaError("%s failed: %s (%s)" % (name, ERRORS.get(rc, rc), load().dasa_last_error().decode()))
    return rc


ROUTES = ("skinny", "pair", "pair_mn", "pair_mn_splitk", "pair_grouped", "tc_single", "simt_tf32_mode", "simt_misaligned",
          "simt_fp32", "pair_f16")


def gemm_route_counts(reset=False):
    """{route name: number of dasa_gemm calls it took} since the last reset (include/dasa_b200.h DASA_ROUTE_*)."""
    buf = (ctypes.c_int64 * len(ROUTES))()
    load().dasa_debug_gemm_route_counts(buf, len(ROUTES), int(bool(reset)))
    return dict(zip(ROUTES, [int(x) for x in buf]))
