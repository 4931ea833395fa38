from .byteTrack.byteTracker import BYTETracker
