PML_LIBRARY=build/libpml_HC1_R200.so python -m pytest tests/test_gpu_parity.py -m gpu -q -k "default_training or golden" 2>&1 | tail -2
for l in HC0_R255 HC1_R255 HC1_R200; do
  for th in 0 96 64; do
    echo "$l PML_TH=$th"
    PML_TH=$th sh tools/ab_libs.sh build/libpml_$l.so | tail -1
  done
done
