OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q > $OUT/r02n2b_pytest.log 2>&1; echo "exit $?" >> $OUT/r02n2b_pytest.log
(cd /tmp && for g in 1 2; do MYTRIM_GPUS=$g MYTRIM_SEED=39172 MYTRIM_TIMING=1 MYTRIM_UO2_CHUNK=16384 timeout 600 $GRAFT_REPO_ROOT/build/apps/mytrim_uo2 uo2out$g 10 0.1 65536 2>&1 | grep workload; done; cmp uo2out1.Erec uo2out2.Erec && cmp uo2out1.dist uo2out2.dist && echo "outputs identical for 1 and 2 GPUs"; wc -l uo2out1.Erec) > $OUT/r02n2b_uo2_app.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r02n2b_bench.json 2> $OUT/r02n2b_bench.err; echo "bench exit $?" >> $OUT/r02n2b_bench.err
tail -5 $OUT/r02n2b_pytest.log; cat $OUT/r02n2b_uo2_app.log; cut -c1-200 $OUT/r02n2b_bench.json
