#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
bash tools/gpu_run14.sh
DUCOSY_PDL=0 bash tools/gpu_prof2.sh
