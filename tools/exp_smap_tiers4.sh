cd /root/repo
run() { echo "== $*"; env "$@" timeout 120 python tools/prof_smap.py 65536 128 0 0 3 1.0 2>&1 | tail -1; }
run TAG=base
run TAG=t54_72_110 SMLVM_SMAP_TIERS=54,72,110
run TAG=t54_80_110 SMLVM_SMAP_TIERS=54,80,110
run TAG=t54_66_110 SMLVM_SMAP_TIERS=54,66,110
run TAG=t48_72_110 SMLVM_SMAP_TIERS=48,72,110
runb() { echo "== $*"; env "$@" timeout 120 python tools/prof_smap.py 65536 128 1 64 3 1.0 2>&1 | tail -1; }
runb TAG=base
runb TAG=t72_88_110 SMLVM_SMAP_TIERS=72,88,110
runb TAG=t72_96_110 SMLVM_SMAP_TIERS=72,96,110
runb TAG=t64_88_110 SMLVM_SMAP_TIERS=64,88,110
runb TAG=t72_84_110 SMLVM_SMAP_TIERS=72,84,110
