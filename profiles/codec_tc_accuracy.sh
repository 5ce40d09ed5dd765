# encoder error vs oracle for the accumulator-chain policies of codec_tc.cu, and the FMA kernel
for mode in "LADIFF_CODEC_SIMT=1" "LADIFF_CODEC_TC_NHI=1" "LADIFF_CODEC_TC_NHI=3" "X=1"; do
  echo "== $mode"
  env $mode timeout 200 python -m pytest tests/test_parity_gpu.py -q -k "cond_codec_codes" 2>&1 | tail -2
  python -c "
import json; d=json.load(open('gpurun_out/parity_report.json'))
for k,v in d.items():
    if k.startswith('cond_encoder'): print(k, v)"
done
