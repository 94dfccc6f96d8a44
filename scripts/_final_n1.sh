for c in 1 0 2 3 4; do
  python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_r02f_cfg$c.json 2> gpurun_out/bench_r02f_cfg$c.err; echo "cfg$c rc=$?"
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02f_reference_arm.json 2> gpurun_out/bench_r02f_reference_arm.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
