#!/bin/bash
# Build the product library and the phase-timer profiling library (tools/phase_prof.py) side by side.
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build(force=True)" > /tmp/build_prod.log 2>&1 &
P1=$!
python tools/phase_prof.py build > /tmp/build_prof.log 2>&1 &
P2=$!
wait $P1; R1=$?; wait $P2; R2=$?
grep -i " error\|failed" /tmp/build_prod.log /tmp/build_prof.log
echo "build rc: product=$R1 prof=$R2"
