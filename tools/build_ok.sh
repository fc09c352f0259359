#!/bin/bash
# build the library; non-zero exit (and no GPU run) when nvcc fails
set -o pipefail
python "$(dirname "$0")/../combo-avs_b200/build.py" "$@" 2>&1 | grep -E "error|libmsda_b200.so$|RuntimeError" ; test -f "$(dirname "$0")/../combo-avs_b200/libmsda_b200.so"
