from .byteTrack.byteTracker import BYTETracker, BYTETrackerPy
