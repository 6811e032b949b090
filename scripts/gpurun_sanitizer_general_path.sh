set -x
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck python scripts/gj_ncu_drive.py 700 > gpurun_out/r02_sanitizer_memcheck_gj.log 2>&1
tail -4 gpurun_out/r02_sanitizer_memcheck_gj.log
timeout 200 compute-sanitizer --tool racecheck python scripts/gj_ncu_drive.py 700 > gpurun_out/r02_sanitizer_racecheck_gj.log 2>&1
tail -6 gpurun_out/r02_sanitizer_racecheck_gj.log
