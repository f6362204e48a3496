import os as _os
# multi-rank runs: eager CUDA module loading (lazy loading of a kernel variant while a peer-waiting NCCL kernel is in flight dead-locks;
# see bench.py).  Effective only if the CUDA context does not exist yet.
if int(_os.environ.get("WORLD_SIZE", "1")) > 1:
    _os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
