run() { python bench.py --steps 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', d['value'], d['ms_per_step'], d['kernels_ms']['pvq_stage(gather+bands+scatter)'], d['kernels_ms']['k_pvq_luma_intra(wavefront)'])"; }
run base
DAALA_B200_SMALL_WAVE=0,0,0 run small_off
DAALA_B200_SMALL_WAVE=32768,65536,0 run small_big
DAALA_B200_SMALL_WAVE=2048,4096,0 run small_small
DAALA_B200_ONE_CHAIN_STREAM=1 run one_stream
