#!/bin/sh
# SASS opcode histogram of the product library: what proves TMA / mbarrier / FP64-FMA code and the absence of the
# reference's recompiled wmma (DMMA) / cp.async (LDGSTS) kernels.   sh profiles/sass_histogram.sh > profiles/r2_sass_histogram.txt
LIB=${1:-lorastencil_b200/lib/liblorastencil_b200.so}
echo "# cuobjdump -sass $LIB  (sm_100a), opcode counts over all kernels"
cuobjdump -sass "$LIB" | grep -oE 'UTMALDG\.[0-9A-Z.]+|UTMASTG\.[0-9A-Z.]+|UTMAPF[.A-Z0-9]*|UBLKCP[.A-Z0-9]*|SYNCS\.[A-Z0-9.]+|DFMA|DMUL|DADD|DMMA[.0-9A-Zx]*|HMMA[.0-9A-Z]*|UTCHMMA|LDGSTS[.A-Z0-9]*|STG\.E\.ENL2\.256|STG\.E\.128|STG\.E\.64|LDS\.128|LDS\.64|STS\.128|SHFL\.[A-Z]+|ATOMG[.A-Z0-9]*|MEMBAR\.[A-Z.]+|STG\.E\.64\.STRONG\.SYS|WARPSYNC[.A-Z]*|BAR\.SYNC[.A-Z]*' | sort | uniq -c | sort -rn
echo
echo "# kernels"
cuobjdump -sass "$LIB" | grep -E "^\s+Function :" | sed 's/^\s*Function : //' | c++filt | sed 's/lora::(anonymous namespace):://'
