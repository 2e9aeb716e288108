#!/bin/bash
# race hunt: the bench parity check of the first blocks while another process time-slices the GPU
(python profiles/contend.py 170 &)
sleep 12
for lib in libfqsk.so libfqsk_prev.so; do for v in default serial profile; do
  [ $lib = libfqsk_prev.so ] && [ $v = serial ] && continue
  echo "== $lib $v"; FQSK_LIB_PATH=$PWD/fqsqueezer_b200/$lib VARIANTS=$v NBLK=4 timeout 120 python profiles/debug_r2g.py parity 2>&1 | tail -2
done; done
