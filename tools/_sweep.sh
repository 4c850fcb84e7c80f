for cfg in "12 64 96" "12 64 128" "12 64 64"; do
  set -- $cfg
  HUF_DEFS="-DHUF_DEC_LOOKUPS=$1 -DHUF_DEC_ROW=$2 -DHUF_DEC_LANES=$3" python huffman-avx512_b200/build.py --force > /dev/null || { echo "build failed $cfg"; continue; }
  echo "lookups=$1 row=$2 lanes=$3 $(python bench.py --no-e2e --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['compress_GBps_per_gpu']), round(d['decompress_GBps_per_gpu']))" 2>&1 | tail -1)"
done
python huffman-avx512_b200/build.py --force > /dev/null
