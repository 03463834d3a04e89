#!/bin/bash
# Build libgpcounts_b200.so for sm_100a (same recipe as __graft_entry__.build()).
set -e
cd "$(dirname "$0")"
python -c "import __graft_entry__ as g; g.build(force=True)"
