# Experiment build ("first thing to try"): fixed-point utility sums by shared-memory atomics on the physics
# side instead of the observers' per-warp walks over the UE bitsets (-DDCB_FX_AGG; not in the default build, never run on
# a GPU yet).  Run here (no GPU needed); the .so travels with the snapshot and is not tracked by git.  Then, on the box:
#   DCB_LIB_PATH=$GRAFT_REPO_ROOT/gpurun_exp_FXAGG.so python -m pytest tests -m gpu -q -x      (parity first)
#   bash scripts/gpu_exp.sh tree gpurun_exp_FXAGG.so                                            (A/B of the headline bench)
cd "$(dirname "$0")/../deepcomp_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false \
    -Xcompiler -fPIC -shared -DDCB_FX_AGG -o ../../gpurun_exp_FXAGG.so *.cu && echo built gpurun_exp_FXAGG.so
