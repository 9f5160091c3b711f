#!/bin/bash
# The driver's GPU test command, verbatim (one pytest process for every -m gpu test), then smoke().
set -x
timeout 600 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
