#!/bin/bash
tools/gpu_ab_env.sh r2p none ACMMP_TILE_ORDER=2 ACMMP_TILE_ORDER=4 ACMMP_TILE_ORDER=8
