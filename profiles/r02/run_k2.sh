#!/bin/bash
# round 2, last GPU session: smoke() + the whole -m gpu suite on the final tree
cd "$(dirname "$0")/../.."
O=gpurun_out/r02k2; mkdir -p $O
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.txt 2>&1; echo "rc=$?" >> $O/smoke.txt
grep "smoke ok\|rc=" $O/smoke.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > $O/pytest_gpu.txt 2>&1; echo "rc=$?" >> $O/pytest_gpu.txt
tail -3 $O/pytest_gpu.txt | cut -c1-200
