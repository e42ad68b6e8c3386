#!/bin/bash
# Round 2, second GPU call (2 GPUs): new tests incl. the 2-rank NCCL ones, driver, bench at N=1 and N=2.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests/test_gpu_round2.py -x -q -m gpu > $O/r02p_gpu_tests_new.txt 2>&1; tail -15 $O/r02p_gpu_tests_new.txt
python -m pytest tests -q -m gpu > $O/r02p_gpu_tests.txt 2>&1; tail -8 $O/r02p_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02p_smoke.txt 2>&1; tail -2 $O/r02p_smoke.txt
python bench.py --steps 20 --warmup 5 > $O/r02p_bench_c4.json 2> $O/r02p_bench_c4.err; tail -c 400 $O/r02p_bench_c4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02p_bench_c4_2gpu.json 2> $O/r02p_bench_c4_2gpu.err; tail -c 600 $O/r02p_bench_c4_2gpu.err
./vision_slam_frontend_b200/vsf_sequence_driver --poses 200 --features 2000 --window 10 --desc-bytes 61 --out $O/driver_one.bin > $O/r02p_driver_1rank.txt 2>&1
for r in 0 1; do ./vision_slam_frontend_b200/vsf_sequence_driver --poses 200 --features 2000 --window 10 --desc-bytes 61 --world 2 --rank $r --device $r --rendezvous /tmp/vsf_id_$$ --out $O/driver_two.bin > $O/r02p_driver_2rank_$r.txt 2>&1 & done; wait
cmp $O/driver_one.bin $O/driver_two.bin && echo "driver: 2-rank bytes == 1-rank bytes" | tee $O/r02p_driver_cmp.txt
cat $O/r02p_driver_1rank.txt $O/r02p_driver_2rank_0.txt $O/r02p_driver_2rank_1.txt
rm -f $O/driver_one.bin $O/driver_two.bin
ls -la $O | tail -12
