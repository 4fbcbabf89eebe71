# evidence run for profiles/: usage: bash scripts/gpu_final_profiles.sh list|cqt|chain
mkdir -p gpurun_out
CMD="python bench.py --clips 288 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
case "$1" in
  list)
    $CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || exit 1
    ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1 ;;
  cqt)
    # the timed chain's launch (all seven octaves in one launch): the 3 warm-up chains are skipped
    ncu --set full --clock-control none --import-source on -k regex:cqt_kernel -s 3 -c 1 -o gpurun_out/prof_final_cqt -f $CMD > gpurun_out/ncu_full_cqt.log 2>&1 ;;
  chain)
    # per chain: stft, tuning, proj, perc, harm, istft, ola, stft, tuning, 7 decimations, tonnetz = 17 launches
    ncu --set full --clock-control none -k "regex:stft_kernel|hpss_harm|hpss_perc|istft_kernel|ola_kernel|proj_kernel|tuning_kernel|decimate2|tonnetz_kernel" -s 51 -c 17 -o gpurun_out/prof_final_chain -f $CMD > gpurun_out/ncu_full_chain.log 2>&1 ;;
esac
echo "exit $?"
