for v in s3w8 s4w8 s3w6; do echo "== $v"; MOC_B200_LIB=/root/repo/tools/libmoc_$v.so python tools/kbench.py --slides 400 2>&1 | grep score_keys; done
echo "== default"; python tools/kbench.py --slides 400 2>&1 | grep score_keys
