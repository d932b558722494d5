"""PPG editing on the device (ppgs/edit/): grid-based time stretching."""
from . import grid  # noqa: F401
