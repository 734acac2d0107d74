#!/bin/bash
for v in default vol3 vol5 vol6 vol8; do
  if [ $v = default ]; then unset B2MTS_LIB; else export B2MTS_LIB=$PWD/mitsuba_b200/libb2mts_$v.so; fi
  echo "$v: $(python scripts/render_once.py smoke 256 512 2>&1 | tail -1)"
done
