#!/bin/bash
MP3B200_LIB=$PWD/lamejs_b200/libmp3b200_stats.so timeout 120 python tools/profile_run.py 10000 1 2>&1 | tail -16
