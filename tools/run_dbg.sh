for a in "GGA 20000 378" "GGA 20000 255" "GGA 40000 377" "B3LYP 30000 377"; do echo "== $a"; timeout 120 python tools/debug_coef.py $a 2>&1 | grep -E "^E|rows differing"; done
