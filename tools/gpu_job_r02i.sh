for r in 255 216 200 184 168; do
  for th in 0 96 64 48; do
    echo "regs $r PML_TH=$th"
    PML_TH=$th sh tools/ab_libs.sh build/libpml_r$r.so | tail -1
  done
done
