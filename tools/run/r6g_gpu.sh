#!/bin/bash
# A/B partner of r6f: the same C2 timing on the build without the unrolled scan
sed -n '/^timeout 60 python - <<.P. 2>&1 | tail -6/,/^P$/p' tools/run/r6f_gpu.sh > /tmp/r6g_part.sh; bash /tmp/r6g_part.sh
