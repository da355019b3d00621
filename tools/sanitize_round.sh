#!/bin/bash
# GPU-side (under gpurun): compute-sanitizer over every execution path of the CURRENT binary; the report names the commit.
#   gpurun -- "AW_GIT_SHA=$(git rev-parse --short HEAD) bash tools/sanitize_round.sh r02"
#   AW_SANITIZE_ONLY=transforms: only the plans that run K2 / K4 / KF; the report goes to gpurun_out/<round>_sanitizer_transforms.txt
R=${1:-r02}
OUT=gpurun_out/${R}_sanitizer${AW_SANITIZE_ONLY:+_$AW_SANITIZE_ONLY}.txt
mkdir -p gpurun_out
{
  echo "# compute-sanitizer over ${AW_SANITIZE_ONLY:-every execution path} (tools/sanitize.py, tools/sanitize_small.py)"
  echo "# commit ${AW_GIT_SHA:-unknown}; library sha256 $(sha256sum airwave_b200/lib/libairwave_cuda.so | cut -c1-16); $(date -u +%FT%TZ)"
  for tool in memcheck synccheck initcheck; do
    echo; echo "## compute-sanitizer --tool $tool python tools/sanitize.py"
    timeout ${AW_SANITIZE_TIMEOUT:-900} compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 | grep -E "^ok|ERROR SUMMARY|=========.*(Invalid|Barrier|Uninitialized|error)" | head -80
  done
  echo; echo "## compute-sanitizer --tool racecheck python tools/sanitize_small.py"
  timeout ${AW_SANITIZE_TIMEOUT:-900} compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | grep -E "^ok|ERROR SUMMARY|RACECHECK SUMMARY|=========.*(hazard|Race|error)" | grep -v "     and \(Write\|Read\) access" | head -120
} > $OUT 2>&1
tail -5 $OUT
