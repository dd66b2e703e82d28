cd /root/repo
run() { echo "== $*"; env "$@" timeout 120 python tools/prof_smap.py 65536 128 0 0 3 1.0 2>&1 | tail -1 | cut -c1-110; }
run TAG=base
run SMLVM_SMAP_TIERS=48,110
run SMLVM_SMAP_TIERS=44,110
run SMLVM_SMAP_TIERS=54,72,110
run SMLVM_SMAP_TIERS=48,72,110
run SMLVM_SMAP_TIERS=44,62,110
run SMLVM_SMAP_TIERS=40,54,72
runb() { echo "== $*"; env "$@" timeout 120 python tools/prof_smap.py 65536 128 1 64 3 1.0 2>&1 | tail -1 | cut -c1-110; }
runb TAG=base
runb SMLVM_SMAP_TIERS=64,110
runb SMLVM_SMAP_TIERS=60,110
runb SMLVM_SMAP_TIERS=72,88,110
runb SMLVM_SMAP_TIERS=64,88,110
runb SMLVM_SMAP_TIERS=56,72,96
runb SMLVM_SMAP_TIERS=60,84,110
