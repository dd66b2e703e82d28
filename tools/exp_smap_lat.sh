cd /root/repo
for r in 148 296 444 592 740 1480 2960; do SMLVM_SMAP_THREADS=128 python tools/prof_smap.py $r 128 0 0 5 1.0 2>&1 | tail -1; done
for r in 148 592; do SMLVM_SMAP_THREADS=128 python tools/prof_smap.py $r 128 1 64 5 1.0 2>&1 | tail -1; done
