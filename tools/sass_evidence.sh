#!/bin/sh
# regenerates profiles/r2_sass_tensor_ops.txt from the built library (no GPU needed): see the python snippet in the profile's header
cuobjdump -sass crowdnav_dsrnn_b200/libcrowdnav_b200.so | grep -E "Function :|UTC|UTMA|UBLKCP|LDTM|STTM|UCGABAR|STG.E.256|LDG.E.256|REDG" | awk '/Function :/{f=$3} !/Function :/{split($0,a," "); for(i in a) if (a[i] ~ /^(UTC|UTMA|UBLKCP|LDTM|STTM|UCGABAR|STG\.E\.256|LDG\.E\.256|REDG)/) c[f" "a[i]]++} END{for(k in c) print k, c[k]}' | sort
