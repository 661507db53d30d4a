"""Host-side tables of a navigation graph for the device-resident environment.

The reference keeps this state in Python dicts: `R2RBatch.paths / .distances` (all-pairs Dijkstra over the connectivity
graph, env.py:182-198, utils.py load_nav_graphs), `buffered_state_dict` (per viewpoint: for every navigable neighbour its
absolute heading / elevation, the view index `pointId` it is seen in and its simulator index, env.py:291-298) and the
per-base-view angle feature tables (utils.py:386-408). Here the same information is flattened ONCE into dense arrays that
live in HBM, so that an environment step is two small kernels instead of Python loops + 5 H2D copies + 1 D2H sync:

  nbr        [n, dmax] int32   neighbour viewpoint per candidate slot (-1 = empty); slot order = candidate order
  nbr_point  [n, dmax] int32   pointId: the view (0..35) the candidate is attached to
  deg        [n]       int32   number of candidates (without END)
  cand_angle [n, dmax, 12, 4]  float32 [sin h, cos h, sin e, cos e] of the candidate relative to each base heading
                               (env.py:303-306: heading = normalized_heading - (viewId % 12) * 30deg)
  view_angle [12, 36, 4]       float32 panorama angle features relative to each base heading (utils.py:386-405)
  agent_angle[36, 4]           float32 angle feature of the agent's own heading / elevation (agent_dg.py:314-317)
  dist       [n, n]    float32 shortest-path length (the reference stores the reward distances in float32, agent_dg.py:897)
  next_hop   [n, n]    int32   candidate slot of the first hop of the shortest path, -1 when already at the goal

Trigonometry is evaluated with math.sin / math.cos on Python floats and rounded to float32, exactly like
utils.angle_feature (utils.py:361-368), so the tables are bit-identical to what the reference concatenates per step.
The shortest-path search follows networkx's Dijkstra (strict improvement, FIFO tie-break on insertion order), which is what
env.py:195-198 calls, so ties resolve the same way.
"""
import heapq
import math
from itertools import count

import numpy as np

R30 = math.radians(30)


def angle4(heading, elevation):
    return [math.sin(heading), math.cos(heading), math.sin(elevation), math.cos(elevation)]


class NavGraph:
    def __init__(self, nbrs, weights, headings, elevations, points, names=None):
        """nbrs[i] = neighbour ids of viewpoint i in candidate order; weights[i][k] = edge length; headings[i][k] /
        elevations[i][k] = absolute ('normalized') heading and elevation of candidate k seen from i; points[i][k] = pointId."""
        self.n = len(nbrs)
        self.names = names or ["vp%05d" % i for i in range(self.n)]
        self.nbrs, self.weights, self.headings, self.elevations, self.points = nbrs, weights, headings, elevations, points
        self.dmax = max(1, max(len(x) for x in nbrs))
        n, d = self.n, self.dmax
        self.nbr = np.full((n, d), -1, np.int32)
        self.nbr_point = np.zeros((n, d), np.int32)
        self.deg = np.zeros(n, np.int32)
        self.cand_angle = np.zeros((n, d, 12, 4), np.float32)
        for i in range(n):
            self.deg[i] = len(nbrs[i])
            for k, j in enumerate(nbrs[i]):
                self.nbr[i, k] = j
                self.nbr_point[i, k] = points[i][k]
                for hb in range(12):
                    self.cand_angle[i, k, hb] = np.array(angle4(headings[i][k] - hb * R30, elevations[i][k]), np.float32)
        self.view_angle = np.zeros((12, 36, 4), np.float32)
        for hb in range(12):
            for ix in range(36):
                self.view_angle[hb, ix] = np.array(angle4((ix % 12) * R30 - hb * R30, (ix // 12 - 1) * R30), np.float32)
        self.agent_angle = np.zeros((36, 4), np.float32)
        for ix in range(36):
            self.agent_angle[ix] = np.array(angle4((ix % 12) * R30, (ix // 12 - 1) * R30), np.float32)
        self.dist64, self.next_hop = self._all_pairs()
        self.dist = self.dist64.astype(np.float32)

    # -------------------------------------------------------------------------------------------- shortest paths
    def _all_pairs(self):
        """All-pairs Dijkstra in the order networkx explores (env.py:195-198): returns float64 distances (inf if
        unreachable) and the candidate slot of the first hop (-1 at the goal or if unreachable)."""
        n = self.n
        dist = np.full((n, n), np.inf, np.float64)
        first = np.full((n, n), -1, np.int32)
        for s in range(n):
            done, seen, hop = {}, {s: 0.0}, {s: -1}
            c = count()
            fringe = [(0.0, next(c), s)]
            while fringe:
                d, _, v = heapq.heappop(fringe)
                if v in done:
                    continue
                done[v] = d
                for k, u in enumerate(self.nbrs[v]):
                    vu = d + self.weights[v][k]
                    if u in done:
                        continue
                    if u not in seen or vu < seen[u]:
                        seen[u] = vu
                        hop[u] = k if v == s else hop[v]
                        heapq.heappush(fringe, (vu, next(c), u))
            for v, d in done.items():
                dist[s, v] = d
                first[s, v] = hop[v]
        return dist, first

    # ------------------------------------------------------------------------------------------------ builders
    @staticmethod
    def from_positions(pos, edges, names=None):
        """pos [n, 3] (x, y, z); edges = iterable of (i, j) in insertion order (undirected). Candidate direction follows the
        simulator's convention: heading is measured from +y towards +x, elevation from the horizontal plane; the candidate
        is attached to the nearest of the 36 discretised views."""
        n = len(pos)
        nbrs, w, hd, el, pt = ([[] for _ in range(n)] for _ in range(5))
        for i, j in edges:
            for a, b in ((i, j), (j, i)):
                if b in nbrs[a]:
                    continue
                dx, dy, dz = (float(pos[b][0] - pos[a][0]), float(pos[b][1] - pos[a][1]), float(pos[b][2] - pos[a][2]))
                heading = math.atan2(dx, dy) % (2 * math.pi)
                elevation = math.atan2(dz, math.hypot(dx, dy))
                level = min(2, max(0, int(round(elevation / R30)) + 1))
                nbrs[a].append(b)
                w[a].append(math.sqrt(dx * dx + dy * dy + dz * dz))
                hd[a].append(heading)
                el[a].append(elevation)
                pt[a].append(level * 12 + int(round(heading / R30)) % 12)
        return NavGraph(nbrs, w, hd, el, pt, names)

    @staticmethod
    def from_connectivity(items):
        """items = the parsed `<scan>_connectivity.json` list (image_id, pose 4x4 row-major, included, unobstructed), edges
        added in the order of utils.load_nav_graphs so networkx and this builder see the same adjacency order."""
        keep = [i for i, it in enumerate(items) if it["included"]]
        idx = {i: k for k, i in enumerate(keep)}
        pos = [[items[i]["pose"][3], items[i]["pose"][7], items[i]["pose"][11]] for i in keep]
        edges = []
        for i in keep:
            for j, conn in enumerate(items[i]["unobstructed"]):
                if conn and items[j]["included"]:
                    edges.append((idx[i], idx[j]))
        return NavGraph.from_positions(pos, edges, [items[i]["image_id"] for i in keep])

    @staticmethod
    def synthetic(n=256, seed=0, max_degree=13, radius=None):
        """Random geometric graph with Matterport-like statistics (mean degree ~4, max 13): viewpoints on a
        jittered multi-floor plan, edges between mutually close viewpoints, made connected by chaining components."""
        rng = np.random.RandomState(seed)
        side = max(2.0, math.sqrt(n) * 2.2)
        pos = np.stack([rng.uniform(0, side, n), rng.uniform(0, side, n), rng.choice([0.0, 0.1, 2.9], n, p=[0.6, 0.25, 0.15]) +
                        rng.normal(0, 0.05, n)], 1)
        radius = radius or 2.9
        d = np.sqrt(((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1))
        deg = np.zeros(n, np.int64)
        edges = []
        order = np.dstack(np.unravel_index(np.argsort(d, axis=None), d.shape))[0]
        for i, j in order:
            if i >= j or d[i, j] > radius:
                continue
            if deg[i] < max_degree and deg[j] < max_degree and rng.rand() < 0.75:
                edges.append((int(i), int(j)))
                deg[i] += 1
                deg[j] += 1
        # connect the components (union-find), nearest pair first, without exceeding max_degree
        parent = list(range(n))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x
        for i, j in edges:
            parent[find(i)] = find(j)
        for i, j in order:
            if i < j and find(i) != find(j) and deg[i] < max_degree and deg[j] < max_degree:
                edges.append((int(i), int(j)))
                deg[i] += 1
                deg[j] += 1
                parent[find(i)] = find(j)
        return NavGraph.from_positions(pos, edges)

    # ------------------------------------------------------------------------------------------------ episodes
    def sample_episodes(self, B, seed=0, min_hops=3, max_hops=7):
        """(start_vp, start_view, goal) int32 arrays: R2R-like tasks, goal 3-7 hops from the start on the shortest path,
        start heading snapped to one of the 12 horizon views (the simulator discretises the dataset's heading)."""
        rng = np.random.RandomState(1000 + seed)
        hops = self.hops()
        start, goal = np.zeros(B, np.int32), np.zeros(B, np.int32)
        for b in range(B):
            for _ in range(1000):
                s = rng.randint(self.n)
                ok = np.nonzero((hops[s] >= min_hops) & (hops[s] <= max_hops))[0]
                if len(ok):
                    start[b], goal[b] = s, ok[rng.randint(len(ok))]
                    break
            else:
                raise RuntimeError("no start/goal pair %d-%d hops apart" % (min_hops, max_hops))
        view = (12 + rng.randint(0, 12, B)).astype(np.int32)
        return start, view, goal

    def hops(self):
        """Number of moves the teacher needs from every start to every goal (follows next_hop); -1 across components."""
        if getattr(self, "_hops", None) is None:
            n = self.n
            h = np.full((n, n), -1, np.int64)
            for g in range(n):
                col = self.dist64[:, g]
                reach = np.nonzero(np.isfinite(col))[0]              # only g's own component (one scan of a multi-scan union)
                # process sources in order of increasing distance to g so the successor is already known
                for s in reach[np.argsort(col[reach], kind="stable")]:
                    k = self.next_hop[s, g]
                    h[s, g] = 0 if k < 0 else h[self.nbr[s, k], g] + 1
            self._hops = h
        return self._hops

    @staticmethod
    def union(graphs, prefixes=None):
        """Disjoint union of several graphs (one per scan, as R2RBatch keeps them: env.py:182-198) in ONE table set, so that a
        batch may mix episodes from different scans: viewpoint ids are offset per graph, distances across graphs are inf and
        next_hop is -1 there. names become '<prefix>_<name>' (the reference's long ids '<scan>_<viewpoint>')."""
        nbrs, w, hd, el, pt, names = [], [], [], [], [], []
        off = 0
        for gi, g in enumerate(graphs):
            pre = (prefixes[gi] if prefixes else "g%d" % gi) + "_"
            for i in range(g.n):
                nbrs.append([j + off for j in g.nbrs[i]])
                w.append(list(g.weights[i])); hd.append(list(g.headings[i])); el.append(list(g.elevations[i]))
                pt.append(list(g.points[i])); names.append(pre + g.names[i])
            off += g.n
        return NavGraph(nbrs, w, hd, el, pt, names)
