cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | wc -l
gcc -std=c99 -D_POSIX_C_SOURCE=200809L -pedantic -Wall -Werror -Iinclude tests/capi_multigpu.c -o /tmp/capi_multigpu -Lncrystal_b200/lib -lncrystal_b200 -lm -Wl,-rpath,$PWD/ncrystal_b200/lib
for nd in 1 2; do /tmp/capi_multigpu "Al_sg225.ncmat;temp=293.15K" $nd 10000000 | tee -a gpurun_out/r2F_capi_multigpu.jsonl; done
NG=2 bash scripts/gpu30.sh 2>&1 | tail -12
timeout 600 python -m pytest tests/test_gpu_api_semantics.py -x -q -k "multi or device" 2>&1 | tail -3
