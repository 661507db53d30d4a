"""TEST INFRASTRUCTURE ONLY (oracle shim). Minimal stand-in for the Matterport3D simulator's pybind
module so that /root/reference/r2r_src/utils.py can be imported (it instantiates a Simulator at import
time, utils.py:704 -> ViewHelper :672-691, and steps it through the 36 discretised views asserting
state.viewIndex == ix). Only the surface utils.new_simulator() (utils.py:370-383) touches is provided.
36 views = 3 elevations x 12 headings, 30 degrees apart (env.py:81-82; MatterSim.cpp:339-363)."""
import math


class _Location:
    viewpointId = ""


class SimState:
    def __init__(self, ix):
        self.viewIndex = ix
        self.heading = (ix % 12) * (math.pi * 2.0 / 12)      # heading_step * headingIncrement (MatterSim.cpp:347-350)
        self.elevation = (ix // 12 - 1) * math.pi / 6.0
        self.location = _Location()
        self.navigableLocations = []


class Simulator:
    def __init__(self):
        self._ix = 0

    def setRenderingEnabled(self, *_): pass
    def setCameraResolution(self, *_): pass
    def setCameraVFOV(self, *_): pass
    def setDiscretizedViewingAngles(self, *_): pass
    def init(self): pass
    def initialize(self): pass

    def newEpisode(self, *_):
        self._ix = 0

    def makeAction(self, *_):
        self._ix = (self._ix + 1) % 36

    def getState(self):
        return SimState(self._ix)
