# signed Plantard kernels at every size against the unsigned ones, with the pipe-levelling knob
# (SPLANT_MAD_NUM/DEN variant builds): bash tools/signed_sweep.sh
L=$PWD/ntt-based-polynomial-multiplier-fpga_b200
bash tools/small_sweep.sh unsigned NTTB200_PLANT_SIGNED=0
bash tools/small_sweep.sh signed_mad1 NTTB200_PLANT_SIGNED=1
for t in mad0 mad12; do
  [ -f $L/libnttb200_$t.so ] && bash tools/small_sweep.sh signed_$t NTTB200_PLANT_SIGNED=1 NTTB200_LIB=$L/libnttb200_$t.so
done
