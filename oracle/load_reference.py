"""TEST INFRASTRUCTURE ONLY — never imported by the product path (dasa_b200/).

Imports the UNMODIFIED reference modules of the agent_dg hot path from /root/reference/r2r_src on CPU so
that (a) the pure-torch restatement in oracle/restated.py can be pinned against the real thing and
(b) golden vectors can be generated (oracle/make_golden.py). /root/reference only exists in the build
container; on the GPU box only the committed fixtures under tests/golden/ and oracle/restated.py travel.

What blocks a plain import and how it is dealt with:
  * param.py parses sys.argv at import and mkdirs snap/<name> in CWD (param.py:200, 252-256)
      -> sys.argv is replaced by the README train command's flags (README.md:82-96) and the import
         happens from a scratch directory.
  * utils.py:7 imports MatterSim and builds a simulator at import (utils.py:704)  -> shims/MatterSim.py
  * env.py:31 loads data/viewpointIds.npy at import                               -> shims/env.py
  * pytorch_transformers / pytorch_pretrained_bert are not installed              -> shims/
"""
import importlib
import os
import sys
import tempfile
import types

REFERENCE_SRC = "/root/reference/r2r_src"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

# README.md:82-96 ("train" command), minus paths that need data.
TRAIN_FLAGS = [
    "--agent_type", "dg", "--adaIn_type", "channel", "--attn", "soft", "--train", "auglistener",
    "--mlWeight_org", "0.4", "--mlWeight_aug", "1.2", "--ab_type", "a", "--a_type", "sigmoid",
    "--d_vl_layers", "3", "--env_drop_stage", "after_adain", "--depth_drop",
    "--use_shift", "--shift_kernel_size", "5",
    "--warm_steps", "1000", "--decay_intervals", "2000", "--decay_start", "4000", "--lr_decay", "0.2",
    "--use_lr_scheduler", "--angleFeatSize", "128", "--accumulateGrad", "--featdropout", "0.4",
    "--feedback", "sample", "--subout", "max", "--optim", "rms", "--lr", "0.0001",
    "--maxAction", "35", "--encoderType", "Dic", "--batchSize", "20",
    "--include_vision", "True", "--use_dropout_vision", "True",
    "--d_enc_hidden_size", "1024", "--critic_dim", "1024", "--name", "oracle_scratch",
]

_cached = None


def available():
    return os.path.isdir(REFERENCE_SRC)


def load(extra_flags=()):
    """Returns a namespace with the reference's `args`, `model`, `r2rmodel`, `vilmodel`, `agent_dg`, `utils`."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_SRC)
    scratch = tempfile.mkdtemp(prefix="dasa_oracle_")
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = ["train.py"] + TRAIN_FLAGS + list(extra_flags)
    os.chdir(scratch)
    # shims must shadow the same-named reference files / missing packages
    sys.path.insert(0, REFERENCE_SRC)
    sys.path.insert(0, _SHIMS)
    try:
        ns = types.SimpleNamespace()
        ns.param = importlib.import_module("param")
        ns.args = ns.param.args
        ns.args.views = 36  # normally a side effect of utils.read_img_features (utils.py:286)
        ns.utils = importlib.import_module("utils")
        ns.model = importlib.import_module("model")
        ns.vilmodel = importlib.import_module("vilmodel")
        ns.r2rmodel = importlib.import_module("r2rmodel")
        ns.agent_dg = importlib.import_module("agent_dg")
        ns.BertConfig = importlib.import_module("pytorch_transformers").BertConfig
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
    _cached = ns
    return ns


if __name__ == "__main__":
    ref = load()
    print("reference imported:", ref.model.__file__, ref.agent_dg.__file__)
