V=kaldi-decoder_b200/lib/variants
for g in 2 3; do tools/sweep_env.sh "160b7_g$g:KD_B200_LIB=$V/libkd_t160b7.so" --groups $g --steps 12 --warmup 4; done
for g in 2 3 4; do tools/sweep_env.sh "128b9_g$g:KD_B200_LIB=$V/libkd_t128b9.so" --groups $g --steps 12 --warmup 4; done
for g in 3 4; do tools/sweep_env.sh "96b12_g$g:KD_B200_LIB=$V/libkd_t96b12.so" --groups $g --steps 12 --warmup 4; done
for g in 2 3; do tools/sweep_env.sh "128b8_g$g:KD_B200_LIB=$V/libkd_t128b8.so" --groups $g --steps 12 --warmup 4; done
