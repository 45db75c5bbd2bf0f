# Instrumented build for scripts/trace_timeline.py (phase timeline of CTA 0): not tracked by git, travels with the snapshot.
cd "$(dirname "$0")/.." && python deepcomp_b200/build.py -DDCB_TRACE -ogpurun_exp_TRACE.so
