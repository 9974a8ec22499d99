#!/bin/bash
# One runner for the GPU-side chores (run under gpurun from the repo root; writes into gpurun_out/):
#   scripts/gpu.sh test                 python -m pytest tests -m gpu
#   scripts/gpu.sh bench [args]         python bench.py [args]  -> gpurun_out/bench.json
#   scripts/gpu.sh launches <workload> [skip] [count]  ncu launch list (gpu__time_duration per launch) of a short bench run
#                                       (keep skip small: ncu intercepts every launch, ~15 ms each even when it is not profiled)
#   scripts/gpu.sh ncu <kernel-regex> <workload> [skip]   one `ncu --set full` capture of a kernel
#   scripts/gpu.sh sanitize             compute-sanitizer memcheck + racecheck over scripts/sanitize_cases.py
set -u
mkdir -p gpurun_out
case "${1:-}" in
test) python -m pytest tests -m gpu -q 2>&1 | tail -30 ;;
bench) shift; python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json ;;
launches) ncu --metrics gpu__time_duration.sum --clock-control none -s "${3:-0}" -c "${4:-140}" --csv --log-file gpurun_out/launches_$2.csv \
    python bench.py --workload "$2" --steps 1 --warmup 1 --sim-steps 200 --skip-cpu --skip-e2e --skip-extra > /dev/null 2>&1 ;;
ncu) ncu --set full --clock-control none --import-source on -k "regex:$2" -s "${4:-50}" -c 1 -o "gpurun_out/prof_$2_$3" \
    python bench.py --workload "$3" --steps 1 --warmup 1 --sim-steps 200 --skip-cpu --skip-e2e --skip-extra > /dev/null 2> gpurun_out/ncu.err; tail -2 gpurun_out/ncu.err ;;
sanitize)
    for tool in memcheck racecheck; do
        compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_cases.py > gpurun_out/sanitize_$tool.log 2>&1
        echo "$tool exit code $?" >> gpurun_out/sanitize_$tool.log
        grep -E "sanitize case|SANITIZE|ERROR SUMMARY|RACECHECK SUMMARY|exit code|=========.*(Error|hazard)" gpurun_out/sanitize_$tool.log | tail -20
    done ;;
*) echo "usage: scripts/gpu.sh test|bench|launches|ncu|sanitize"; exit 2 ;;
esac
