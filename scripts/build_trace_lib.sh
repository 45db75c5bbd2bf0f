# Instrumented build for scripts/trace_timeline.py (clock64 at the phase boundaries, per-CTA wall times).  Run here (no GPU
# needed) before `gpurun -- bash scripts/gpu_final.sh <tag>`; the .so travels with the snapshot and is not tracked by git.
cd "$(dirname "$0")/../deepcomp_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false \
    -Xcompiler -fPIC -shared -DDCB_TRACE -o ../../gpurun_exp_TRACE.so *.cu && echo built gpurun_exp_TRACE.so
