# Two-GPU visit (gpurun --gpus 2): bench at N=2 as the driver launches it, the GPU tests that need two devices, and the
# multi-GPU mytrim_uo2 driver.
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/r02n2_gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r02n2_bench.json 2> $OUT/r02n2_bench.err; echo "bench exit $?" >> $OUT/r02n2_bench.err
timeout 300 python -m pytest tests -m gpu -q -k "allreduce_in_process or chunks_or_gpus" > $OUT/r02n2_pytest.log 2>&1; echo "exit $?" >> $OUT/r02n2_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --workload uo2_fission --no-configs --no-cpu-baseline > $OUT/r02n2_bench_uo2.json 2> $OUT/r02n2_bench_uo2.err; echo "bench exit $?" >> $OUT/r02n2_bench_uo2.err
(cd /tmp && for g in 1 2; do MYTRIM_GPUS=$g MYTRIM_SEED=39172 MYTRIM_TIMING=1 MYTRIM_UO2_CHUNK=16384 timeout 600 $GRAFT_REPO_ROOT/build/apps/mytrim_uo2 uo2out$g 10 0.1 65536 2>&1 | grep workload; done; cmp uo2out1.Erec uo2out2.Erec && cmp uo2out1.dist uo2out2.dist && echo "outputs identical for 1 and 2 GPUs") > $OUT/r02n2_uo2_app.log 2>&1
cat $OUT/r02n2_bench.json | cut -c1-300; tail -2 $OUT/r02n2_bench.err; tail -4 $OUT/r02n2_pytest.log; cat $OUT/r02n2_bench_uo2.json | cut -c1-300; tail -2 $OUT/r02n2_bench_uo2.err; cat $OUT/r02n2_uo2_app.log
