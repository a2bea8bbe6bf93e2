"""Re-export: the Python view of include/sp_flat_batch.h lives in the package (secphase_b200/flatbatch.py)."""
from secphase_b200.flatbatch import CFlatBatch, FlatBatch, _FIELDS  # noqa: F401
