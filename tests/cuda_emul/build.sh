#!/bin/sh
# builds the CPU emulation of the kernels (test infrastructure only)
set -e
cd "$(dirname "$0")"
mkdir -p _build
g++ -std=c++17 -O1 -g -fPIC -shared -Wall -Wno-unknown-pragmas -Wno-unused-variable -Wno-unused-but-set-variable \
    -o _build/libemul.so emul_kernels.cpp
