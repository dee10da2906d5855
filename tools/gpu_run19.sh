#!/bin/bash
# ncu --set full of ONE big layer per forced variant (profiler window = the last forward pass)
export PYTHONUNBUFFERED=1
O=gpurun_out; mkdir -p $O
python -c "import ctypes; ctypes.CDLL('libcudart.so.12'); print('cudart ok')" || find / -name "libcudart.so*" 2>/dev/null | head -3
for V in 0 7; do
  SB_FORCE_VARIANT=$V timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_conv_tc \
    -o $O/layer_512_256_v$V -f python tools/prof_layer.py 512 256 3 64 64 8 1 > $O/ncu_layer_v$V.log 2>&1; echo "v$V rc=$?"; tail -2 $O/ncu_layer_v$V.log
done
SB_FORCE_VARIANT=7 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_conv_tc \
    -o $O/layer_256_128_v7 -f python tools/prof_layer.py 256 128 3 128 128 8 1 > $O/ncu_layer2_v7.log 2>&1; echo "rc=$?"; tail -2 $O/ncu_layer2_v7.log
ls -la $O/*.ncu-rep
