"""CUDA-kernel backend of the CP-CSV drop-in modules (libcpcsv.so behind ctypes)."""
import os

# The step runs up to ~8 independent branches (three discriminators, two generator calls, weight
# re-layout, detached generator forward) on parallel streams / CUDA-graph branches.  With the
# default of 8 hardware work queues, branches that share a queue serialise behind each other's
# waits (measured: one discriminator update ran 6 ms late).  Only effective when set before
# the CUDA context is created, hence at import time.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
