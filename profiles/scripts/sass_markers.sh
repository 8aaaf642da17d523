#!/bin/bash
# SASS evidence that the contractions are tcgen05 / TMEM / TMA: per kernel, the count of the marker mnemonics (runs on the CPU box)
SO=comfyui-sdmatte_b200/libsdmatte_b200.so
cuobjdump -sass $SO | awk '
/Function :/ { fn=$3 }
/UTCHMMA/ { mma[fn]++ } /LDTM/ { ldtm[fn]++ } /STTM/ { sttm[fn]++ } /UTMALDG/ { tma[fn]++ } /UTCBAR/ { bar[fn]++ } / HMMA|\tHMMA/ { legacy[fn]++ } /MUFU.EX2/ { ex2[fn]++ }
END { printf "%-8s %-6s %-6s %-8s %-7s %-6s %-8s %s\n", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "HMMA", "MUFU.EX2", "kernel";
      for (f in mma) printf "%-8d %-6d %-6d %-8d %-7d %-6d %-8d %s\n", mma[f], ldtm[f], sttm[f], tma[f], bar[f], legacy[f], ex2[f], f }' | (read h; echo "$h"; sort -k8 | c++filt)
