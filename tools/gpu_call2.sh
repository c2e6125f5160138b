#!/bin/bash
mkdir -p gpurun_out
echo "== no-tma probe"; ACMMP_NO_TMA=1 timeout 120 python tools/debug_probe.py > gpurun_out/dbg_notma.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/dbg_notma.log
echo "== tma probe under memcheck"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/debug_probe.py > gpurun_out/dbg_sanitizer.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/dbg_sanitizer.log | head -60
