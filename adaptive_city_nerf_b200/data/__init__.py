"""Ray-batch producers (reference: data/task_dataset.py) -- the part that runs per ray."""
from .task_binning import TaskGrid, route_and_bin  # noqa: F401
