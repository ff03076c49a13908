"""Warning / error classes of the hot path.

Same names and roles as qinfer/_exceptions.py:54-76.  When the real QInfer is
importable its classes are re-used, so that user code filtering on
``qinfer.ApproximationWarning`` etc. keeps working after the switch.
"""
try:  # pragma: no cover - QInfer is not installed on the GPU box
    from qinfer._exceptions import ApproximationWarning, ResamplerWarning, ResamplerError
except Exception:
    class ApproximationWarning(RuntimeWarning):
        """Raised when a numerical approximation fails in a way that may violate
        assumptions, for instance when a resampling step fails."""

    class ResamplerWarning(RuntimeWarning):
        """Warning raised in response to events within resampling steps."""

    class ResamplerError(RuntimeError):
        """Error raised when a resampler encounters an unrecoverable condition."""

        def __init__(self, msg, cause=None):
            super(ResamplerError, self).__init__(msg)
            self._cause = cause


class UnsupportedModelError(TypeError):
    """The model is not one of the built-in families the CUDA kernels implement.
    There is deliberately no CPU fallback."""
