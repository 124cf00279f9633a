# Counts the Blackwell-specific SASS instructions in the built library and names the kernel each sits in (no GPU needed).
# usage: bash scripts/sass_evidence.sh > profiles/rNN_sass_tcgen05_tma.txt
set -e
SO=multimodalgame_b200/libmmg_b200.so
T=$(mktemp); cuobjdump -sass $SO > $T
PAT='UTMALDG[.A-Z0-9]*|UTCHMMA|LDTM[.a-z0-9]*|LDGMC[.A-Za-z0-9]*|UTCBAR|UBLKCP[.A-Z]*'
echo "# cuobjdump -sass $SO (sm_100a): counts of the Blackwell-specific instructions in the shipped library"
grep -oE "\b(UTMALDG[.A-Z0-9]*|UTCHMMA|LDTM[.a-z0-9]*|LDGMC[.A-Za-z0-9.]*|UBLKCP[.A-Z]*|UTCBAR|UTCATOMSWS[.A-Z_]*|SYNCS[.A-Z0-9_]*|LDGSTS[.A-Z0-9]*|FFMA2|ACQBULK)\b" $T | sort | uniq -c | sort -rn
echo; echo "# which kernel each one sits in"
awk -v pat="$PAT" '/Function :/ {f=$3} $0 ~ pat {match($0, pat); k=f" "substr($0,RSTART,RLENGTH); c[k]++} END {for (k in c) print c[k], k}' $T | sort -k2 | c++filt | cut -c1-200
echo; echo "# sample lines"
for m in UTMALDG UTCHMMA LDTM LDGMC UBLKCP; do grep -m1 "$m" $T; done
rm -f $T
