#!/bin/bash
O=gpurun_out/r02p; mkdir -p $O
timeout 300 python scripts/prof_e2e.py > $O/e2e_phases.log 2>&1; cat $O/e2e_phases.log
