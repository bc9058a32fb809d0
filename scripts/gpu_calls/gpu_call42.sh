#!/bin/bash
bash scripts/ncu_cases.sh r2 conv_fwd_cfg3 conv_dgrad_cfg3
