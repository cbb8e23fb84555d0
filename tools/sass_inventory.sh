#!/usr/bin/env bash
# Per-kernel counts of the Blackwell-only SASS mnemonics in libedgeline_b200.so (run anywhere cuobjdump exists; no GPU needed):
#   UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy, SYNCS = mbarrier, UTCBAR = tcgen05.commit
set -euo pipefail
LIB=${1:-edge_yolo_b200/libedgeline_b200.so}
cuobjdump -sass "$LIB" | c++filt | awk '
  /Function :/ { fn=$0; sub(/.*Function : /, "", fn); sub(/\(.*/, "", fn) }
  /UTCHMMA/ {c[fn,"UTCHMMA"]++} /LDTM/ {c[fn,"LDTM"]++} /UTMALDG/ {c[fn,"UTMALDG"]++} /UTMASTG/ {c[fn,"UTMASTG"]++} /UBLKCP/ {c[fn,"UBLKCP"]++}
  /SYNCS/ {c[fn,"SYNCS"]++} /UTCBAR/ {c[fn,"UTCBAR"]++} /[^C]HMMA/ {c[fn,"HMMA"]++} /UTCATOMSWS|UTCALLOC/ {c[fn,"TMEM_ALLOC"]++}
  /Function :/ { fns[fn]=1 }
  END { n=split("UTCHMMA LDTM UTMALDG UTMASTG UBLKCP SYNCS UTCBAR TMEM_ALLOC HMMA", k, " ");
        printf "%-72s", "kernel"; for (i=1;i<=n;i++) printf "%10s", k[i]; printf "\n";
        for (f in fns) { tot=0; for (i=1;i<=n;i++) tot+=c[f,k[i]]; if (tot>0) { printf "%-72s", substr(f,1,72); for (i=1;i<=n;i++) printf "%10d", c[f,k[i]]; printf "\n" } } }' | (read -r h; echo "$h"; sort)
