"""visual_sgraphs_b200 — B200-native ORB feature front-end (extractor + Hamming matcher) for vS-Graphs.

The compute lives in csrc/ (CUDA, sm_100a) behind the C ABI declared in include/vsg_cuda.h
(libvsg_cuda.so).  This Python package is a thin ctypes mirror of that ABI used by the tests and
bench.py; the C++ drop-in classes VS_GRAPHS::ORBextractor / ORBmatcher are under shim/.
"""
__version__ = "0.1.0"
