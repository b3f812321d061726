set -x
ncu --set full --clock-control none --import-source on -k regex:'k_bin_reduce|k_sort_scatter|k_bin_index|k_sort_hist' -s 12 -c 6 -o gpurun_out/r01_bin_full python scripts/prof_bin.py > /dev/null 2>&1
ls -la gpurun_out/
