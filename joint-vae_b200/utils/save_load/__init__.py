from .recorders import LossRecorder      # noqa: F401
