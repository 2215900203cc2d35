#!/bin/bash
# build libpifpaf_b200.so in-tree (sm_100a) and print what was built
set -e
cd "$(dirname "$0")/../openpifpaf_b200/csrc"
make 2>&1 | grep -i -E "error|Error" -A3 || true
ls -la --time-style=full-iso libpifpaf_b200.so net.o | awk '{print $6, $7, $9}'
