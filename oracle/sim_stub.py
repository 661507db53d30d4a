"""TEST INFRASTRUCTURE ONLY. A graph-aware stand-in for the Matterport3D simulator in discretised-view mode, just enough to
drive the UNMODIFIED reference env.py / agent_dg.py functions in tests/test_oracle_env_vs_reference.py:
  newEpisode / makeAction / getState with viewIndex, heading = step * 2pi/12, elevation in {-30, 0, +30} degrees
  (src/lib/MatterSim.cpp:339-363, 470-490), navigableLocations[0] = the current location, [k + 1] = candidate slot k.
The real simulator lists only the locations visible in the current view; the agent indexes the list with the candidate's
stored `idx` right after turning to the candidate's pointId (agent_dg.py:386-390), so listing every neighbour at a fixed index is
equivalent for the rollout."""
import math

INC = math.pi * 2.0 / 12


class _Loc:
    def __init__(self, viewpointId, rel_heading=0.0, rel_elevation=0.0):
        self.viewpointId, self.rel_heading, self.rel_elevation = viewpointId, rel_heading, rel_elevation


class _State:
    pass


class GraphSim:
    def __init__(self, scan, names, nbrs):
        self.scan, self.names, self.nbrs = scan, names, nbrs
        self.index = {n: i for i, n in enumerate(names)}
        self.vp, self.hstep, self.level = 0, 0, 1

    def newEpisode(self, scanId, viewpointId, heading, elevation):
        self.vp = self.index[viewpointId]
        self.hstep = int(round((heading % (2 * math.pi)) / INC)) % 12
        self.level = 0 if elevation < -INC / 2 else (2 if elevation > INC / 2 else 1)

    def makeAction(self, index, heading, elevation):
        if index > 0:
            self.vp = self.nbrs[self.vp][index - 1]
        if heading > 0:
            self.hstep = (self.hstep + 1) % 12
        elif heading < 0:
            self.hstep = (self.hstep - 1) % 12
        if elevation > 0:
            self.level = min(2, self.level + 1)
        elif elevation < 0:
            self.level = max(0, self.level - 1)

    def getState(self):
        s = _State()
        s.scanId = self.scan
        s.location = _Loc(self.names[self.vp])
        s.viewIndex = self.level * 12 + self.hstep
        s.heading = self.hstep * INC
        s.elevation = (self.level - 1) * (math.pi / 6.0)
        s.navigableLocations = [s.location] + [_Loc(self.names[j]) for j in self.nbrs[self.vp]]
        return s
