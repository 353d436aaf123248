mkdir -p gpurun_out
FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200_tl.so timeout 120 python profiles/timeline.py > gpurun_out/r49_timeline.txt 2>&1
FREEFINE_B200_LIB=$PWD/freefine_b200/lib/libfreefine_b200_tlko.so timeout 120 python profiles/timeline.py > gpurun_out/r49_timeline_ko.txt 2>&1
tail -3 gpurun_out/r49_timeline.txt; tail -3 gpurun_out/r49_timeline_ko.txt
