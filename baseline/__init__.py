"""Tuned CPU arm of the benchmark (bench.py --impl reference / cpu_baseline).  Not the product, not the checker."""
