#!/bin/bash
# The driver's GPU test command, verbatim (one pytest process for every -m gpu test).
set -x
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
