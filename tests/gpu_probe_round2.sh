#!/bin/bash
# Development probe (gpurun): every listed test id in its OWN process under a hard timeout, short tracebacks.
mkdir -p gpurun_out
run() {  # name, timeout, pytest args...
  local name=$1 to=$2; shift 2
  echo "=== $name"
  timeout -s KILL $to python -m pytest "$@" -q --tb=short -x 2>&1 | tail -${TAILN:-30}
  echo "--- rc=${PIPESTATUS[0]}"
}
export FDX_ATTN_UNFUSED=${FDX_ATTN_UNFUSED:-}
[ -z "$FDX_ATTN_UNFUSED" ] && unset FDX_ATTN_UNFUSED
for t in "$@"; do
  run "$t" ${PROBE_TIMEOUT:-150} "$t"
done
