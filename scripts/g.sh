#!/bin/bash
# build everything, then run a script on the GPU box:  scripts/g.sh [--gpus N] <timeout> <script>
set -e
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" 2>&1 | grep -v "deprecated" || true
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
/usr/local/graft/bin/gpurun $GP --timeout $1 -- "bash $2"
