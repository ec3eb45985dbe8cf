for cfg in "4 4" "8 4" "2 2" "4 2" "1 1"; do
  set -- $cfg
  MPSB_NVCC_FLAGS="-DMPSB_QD_WU=$1 -DMPSB_QD_TILES=$2" python -c "import __graft_entry__ as g; g.build(force=True)" > /dev/null 2>&1
  python bench.py --steps 4 --warmup 3 --no-cpu --no-idmrg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('WU/TILES', '$cfg', d['value'], d['phases_ms_per_step']['qr'])"
done
python -c "import __graft_entry__ as g; g.build(force=True)" > /dev/null 2>&1
