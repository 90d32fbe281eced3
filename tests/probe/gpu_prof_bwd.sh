#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
VAESNE_TC_PROFILE=1 python vaesne-dev_b200/build.py --force > /dev/null 2>&1
ONLY_TIME=1 TC_PROF=1 PS=${PS:-0.0,0.1} python tests/probe/attn_tc_check.py 2>&1 | tail -16
