#!/bin/bash
timeout 300 python tools/gemm_epi_sweep.py 2>&1 | tail -40
