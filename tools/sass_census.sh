#!/bin/bash
# SASS mnemonic census of libvqb200.so (what proves a Blackwell-native kernel: B200_PROFILING.md)
SO=${1:-vector_quantization_b200/libvqb200.so}
echo "# cuobjdump -sass $SO | mnemonic counts ($(date -u +%F))"
cuobjdump -sass "$SO" > /tmp/vqb_sass.txt
for m in UTCHMMA UTCQMMA LDTM STTM UTMALDG UTMASTG UBLKCP " HMMA" " HGMMA" SYNCS.ARRIVE SYNCS.PHASECHK UTCBAR ELECT ACQBULK LDGSTS "RED.E.ADD.F32" "REDG" "LDG.E.STRONG.SYS" "STG.E.STRONG.SYS" "MEMBAR.SC.SYS" "MEMBAR.ALL.SYS" FMNMX3 "FMNMX "; do
  printf "%-20s %s\n" "$m" "$(grep -c -- "$m" /tmp/vqb_sass.txt)"
done
echo "# kernels"
cuobjdump -sass "$SO" | grep "Function :" | sed 's/.*Function : //' | sort | uniq -c | sort -rn | awk '{print $2}' | c++filt | sed 's/(.*//' | sort | uniq -c | sort -rn | head -60
echo "# arch"; grep -m1 "arch =" /tmp/vqb_sass.txt
