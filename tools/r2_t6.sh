for S in 0 5 7 9 11 13 14 15 17 20; do echo "db_splits=$S"; timeout 200 python tools/shard_emulate.py --workload flat --world 8 --option db_splits=$S --steps 8 2>&1 | grep "scan(shard)"; done
for S in 0 3 4 5 6 7 9; do echo "c5 db_splits=$S"; timeout 200 python tools/shard_emulate.py --workload c5 --world 8 --option db_splits=$S --steps 8 2>&1 | grep "scan(shard)"; done
for W in 2 4; do echo "flat world=$W auto"; timeout 200 python tools/shard_emulate.py --workload flat --world $W --steps 8 2>&1 | grep "scan(shard)"; done
