mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_step.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/sanitize_step.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r02_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python scripts/sanitize_step.py > gpurun_out/r02_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r02_sanitizer_synccheck.log
