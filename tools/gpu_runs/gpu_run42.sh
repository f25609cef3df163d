#!/bin/bash
timeout 120 python tools/fused_dbg.py 256 2>&1 | tail -12
