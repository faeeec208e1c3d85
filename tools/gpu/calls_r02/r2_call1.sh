#!/bin/bash
# Round 2, GPU call 1: parity at the benchmarked sizes, the sliced walk (first run on a GPU), probes.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_gpu.txt 2>&1
nproc >> $O/r2a_gpu.txt
# 1. the whole GPU suite, new at-size tests included
timeout 1500 python -m pytest tests -m gpu -q -rs --durations=15 > $O/r2a_pytest.log 2>&1
echo "pytest rc $?" >> $O/r2a_pytest.log
# 2. sliced walk: frames identical across drivers 0,1,2,3?
SOLR_B200_LIB=$PWD/sol-r_b200/csrc/libvar_slice.so SOLR_B200_NVCC_FLAGS=-DWITH_TRACE_SLICE timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "drivers_produce" > $O/r2a_slice_test.log 2>&1
echo "slice test rc $?" >> $O/r2a_slice_test.log
# 3. sweeps on config 2: staged default, then sliced with different rounds / slice counts
( SOLR_B200_LIB=$PWD/sol-r_b200/csrc/libvar_slice.so SOLR_OPT6=1 timeout 200 python tools/gpu/gpu_option_sweep.py 8 300 c2
  SOLR_B200_LIB=$PWD/sol-r_b200/csrc/libvar_slice.so SOLR_OPT6=3 timeout 300 python tools/gpu/gpu_option_sweep.py 10 8,12,16,24,32,48 c2
  SOLR_B200_LIB=$PWD/sol-r_b200/csrc/libvar_slice.so SOLR_OPT6=3 timeout 300 python tools/gpu/gpu_option_sweep.py 11 3,4,8,12 c2 ) > $O/r2a_slice_sweep.log 2>&1
# 4. probes: OpenCL on the box (BASELINE.json config 1 names PoCL), FP32 FMA peak
( which clinfo && clinfo | head -40; ls -la /etc/OpenCL/vendors 2>&1; ldconfig -p | grep -i -E "opencl|pocl"; python -c "import pyopencl" 2>&1 | tail -1;
  find / -name "libOpenCL*" -not -path "/proc/*" 2>/dev/null | head; find / -iname "*pocl*" -not -path "/proc/*" 2>/dev/null | head ) > $O/r2a_opencl_probe.txt 2>&1
python - > $O/r2a_fp32_peak.txt 2>&1 <<'PY'
import sys, os, ctypes as C
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
from _solr_b200_import import solr_b200
from solr_b200 import engine
lib = engine.load()
for k in range(3):
    t, m = C.c_float(), C.c_float()
    rc = lib.b200_measure_fp32_peak(C.byref(t), C.byref(m))
    print("fp32 peak rc %d: %.2f TFLOP/s at %.0f MHz" % (rc, t.value, m.value))
PY
tail -3 $O/r2a_pytest.log; tail -3 $O/r2a_slice_test.log; cat $O/r2a_slice_sweep.log | tail -30; cat $O/r2a_fp32_peak.txt
