cd /root/repo
export SMLVM_SMAP_IMPL=cta
run() { echo "== $*"; env "$@" python tools/prof_smap.py 32768 128 0 0 3 1.0 2>&1 | tail -1; }
run TAG=base
run TAG=t54_110 SMLVM_SMAP_TIERS=54,110
run TAG=t54_72 SMLVM_SMAP_TIERS=54,72
run TAG=t52_72 SMLVM_SMAP_TIERS=52,72
run TAG=t54_72_gf SMLVM_SMAP_TIERS=54,72 SMLVM_SMAP_GRID_FACTOR=6
runb() { echo "== $*"; env "$@" python tools/prof_smap.py 32768 128 1 64 3 1.0 2>&1 | tail -1; }
runb TAG=base
runb TAG=t64_96 SMLVM_SMAP_TIERS=64,96
runb TAG=t80_110 SMLVM_SMAP_TIERS=80,110
runb TAG=t72_96 SMLVM_SMAP_TIERS=72,96
runb TAG=t72_128thr SMLVM_SMAP_THREADS=128
