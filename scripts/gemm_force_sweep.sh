cd "$(dirname "$0")/.." || exit 1
run() { # M N K mode res
  echo "== M=$1 N=$2 K=$3 mode=$4 res=$5"
  python scripts/one_gemm.py $1 $2 $3 $4 $5 10
  for f in "1,32" "1,64" "1,96" "1,128" "1,160" "1,192" "1,256" "2,64" "2,96" "2,128" "2,160" "2,192" "1,64,1" "1,128,1" "1,192,1" "1,256,1"; do
    CAP4D_GEMM_FORCE=$f python scripts/one_gemm.py $1 $2 $3 $4 $5 10 2>/dev/null | tail -n 1
  done
}
run 327680 960 320 1 0
run 327680 2560 320 2 0
run 327680 320 1280 1 1
run 327680 320 320 0 1
run 327680 320 320 0 0
