# Kernel A/B runs: every kimimaro_b200/_variants/<name>.so is another build of the same sources
# (see DESIGN.md); B2T_LIB selects it.  Usage: bash scripts/run_variants.sh <script.py> <name>...
script=$1; shift
for rep in 1 2; do
  for v in "$@"; do
    echo "== $v (rep $rep)"
    B2T_LIB=$PWD/kimimaro_b200/_variants/$v.so timeout 300 python $script 2>&1 | tail -3
  done
done
