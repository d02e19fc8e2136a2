#!/bin/bash
for v in lamejs_b200/libmp3b200_st*.so; do echo "== $v"; MP3B200_LIB=$PWD/$v timeout 120 python tools/profile_run.py 10000 1 2>&1 | grep -E "bs1|timings"; done
