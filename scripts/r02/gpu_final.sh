#!/bin/bash
# End-of-round pass on one GPU: parity tests, smoke, bench (ours + reference arm), ncu launch list, ncu --set full of
# the small kernels and of finalize_kernel.
set +e
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > $O/fin_pytest.log 2>&1
echo "pytest exit $?" >> $O/fin_pytest.log; tail -3 $O/fin_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/fin_smoke.log 2>&1; echo "smoke exit $?" >> $O/fin_smoke.log; tail -2 $O/fin_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/fin_bench.json 2> $O/fin_bench.err; echo "bench exit $?" >> $O/fin_bench.err
tail -2 $O/fin_bench.err; head -c 700 $O/fin_bench.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/fin_bench_reference.json 2> $O/fin_bench_reference.err; echo "ref exit $?" >> $O/fin_bench_reference.err
head -c 400 $O/fin_bench_reference.json; echo
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu --no-decode > $O/fin_ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/fin_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-decode > $O/fin_ncu_launches.log 2>&1
tail -1 $O/fin_ncu_launches.log | cut -c1-200
timeout 600 ncu --set full --clock-control none -k regex:"half_sums_chain|tables_from_halves|affine_kernel|finalize_kernel" -c 4 -f -o $O/fin_small python bench.py --steps 1 --warmup 0 --no-cpu --no-decode > $O/fin_ncu2.log 2>&1
tail -1 $O/fin_ncu2.log | cut -c1-200
