#!/usr/bin/env bash
# e2e (pgw_step_host on page-locked buffers) against the observation staging and the action-fetch stagger
for so in ${SWEEP_STAGE:-0 1}; do for st in ${SWEEP_STAGGER:-0 5 10 20}; do
  echo "PGW_STAGE_OBS=$so PGW_STAGGER_CYCLES_PER_ROW=$st"
  PGW_STAGE_OBS=$so PGW_STAGGER_CYCLES_PER_ROW=$st python tools/e2e_probe2.py 2>&1 | head -1 | cut -c1-200
done; done
