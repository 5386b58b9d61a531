#!/bin/bash
# Round-2 closing sanity job (1 GPU, < 1 min): the library as finally built loads and runs — smoke() + two fast test files.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out; mkdir -p $O; rm -f $O/r2fin5_*
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2fin5_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2fin5_status.log
timeout 60 python -m pytest tests/test_optim_gpu.py tests/test_graph_gpu.py -m gpu -q > $O/r2fin5_tests.log 2>&1; echo "tests rc=$?" >> $O/r2fin5_status.log
cat $O/r2fin5_status.log; tail -1 $O/r2fin5_smoke.log | cut -c1-200; tail -1 $O/r2fin5_tests.log | cut -c1-200
