echo "== with saves"; python scripts/phase_timing.py libmmg_dbg.so 2>&1 | tail -10
echo "== without global saves"; python scripts/phase_timing.py libmmg_nosave.so 2>&1 | tail -10
