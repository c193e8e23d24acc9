python tools/latency_variants.py 1000000 1
python tools/latency_variants.py 1000000 4
python tools/latency_variants.py 25000 1
