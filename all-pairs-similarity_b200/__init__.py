"""B200-native inverted-index all-pairs similarity scoring (drop-in for the IndexingWorkerActor path
of mcgill-cpslab/all-pairs-similarity).  CUDA kernels + C ABI in csrc/, host mirror of the
reference's message protocol in messages.py / worker.py, shard dispatcher in dispatcher.py."""
from . import etl, loadgen, messages, native, synth, worker  # noqa: F401

__all__ = ["native", "synth", "messages", "worker", "loadgen"]
