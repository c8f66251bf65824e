# Copies what round2_refresh.sh <tag> brought back in gpurun_out/ into profiles/r02 (summaries and hot lines are made here
# from the .ncu-rep files).  Usage: bash profiles/scripts/install_refresh.sh <tag>
tag=${1:-v2}; O=gpurun_out; P=profiles/r02
cp $O/traffic.json $P/traffic.json
for w in amars_lw:524288 amars_sw:512000 lbl:800000; do n=${w%%:*}; s=${w##*:}
  python profiles/scripts/ncu_summary.py $O/prof_${n}_$tag.ncu-rep $s > $P/ncu_full_${n}_${tag}_summary.txt
  python profiles/scripts/ncu_hotlines.py $O/prof_${n}_$tag.ncu-rep 25 > $P/hotlines_${n}_$tag.txt
done
(for k in split_prefix split_layer team_adding; do python profiles/scripts/ncu_summary.py $O/prof_gcm_$tag.ncu-rep 75776 $k; echo; done) > $P/ncu_full_gcm_${tag}_summary.txt
python profiles/scripts/ncu_summary.py $O/prof_amars_lw_$tag.ncu-rep 524288 band_optics > $P/ncu_full_optics_${tag}_summary.txt
python profiles/scripts/ncu_hotlines.py $O/prof_gcm_$tag.ncu-rep 30 > $P/hotlines_gcm_$tag.txt
for f in bench_lw bench_amars_sw bench_gcm bench_lbl bench_lblwave_1gpu bench_reference; do cp $O/${f}_$tag.json $P/; done
cp $O/launches_lw_$tag.csv $O/launches_gcm_$tag.csv $O/parity_report_$tag.txt $O/latency_$tag.txt $P/
cp $P/parity_report_$tag.txt $P/parity_report.txt
