# usage: bash tools/_ncu_capture.sh NAME "KERNEL_REGEX" SKIP COUNT UNITS -- command...
# one `ncu --set full` capture condensed ON THE BOX into gpurun_out/r2_NAME_ncu_full.txt and, for launch 0,
# the executed instruction mix gpurun_out/r2_NAME_sass_mix.txt (per UNITS); the report itself stays in /tmp
NAME=$1; REGEX=$2; SKIP=$3; COUNT=$4; UNITS=$5; shift 6
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$REGEX" -s $SKIP -c $COUNT -o /tmp/prof_$NAME -f "$@" > /tmp/ncu_$NAME.log 2>&1
echo "ncu $NAME exit $?"
python tools/ncu_summary.py /tmp/prof_$NAME.ncu-rep > gpurun_out/r2_${NAME}_ncu_full.txt 2>&1
ncu -i /tmp/prof_$NAME.ncu-rep --page source --csv --print-source sass > /tmp/src_$NAME.csv 2>/dev/null
python tools/sass_mix.py /tmp/src_$NAME.csv 0 $UNITS > gpurun_out/r2_${NAME}_sass_mix.txt 2>&1
rm -f /tmp/prof_$NAME.ncu-rep /tmp/src_$NAME.csv
