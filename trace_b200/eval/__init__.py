"""Host-side post-processing of generated ids: the output parser (trace_b200.events) and the three reformatters
the reference runs before its metric scripts (trace/eval/reformat_{dvc,tvg,vhd}.py)."""
from .reformat import reformat_dvc, reformat_tvg, reformat_vhd, result_record  # noqa: F401
