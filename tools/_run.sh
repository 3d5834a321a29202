cp deltarice_b200/libh5deltarice_b200.so /tmp/orig.so
for v in 100 300; do cp tools/build/var/lib_$v.so deltarice_b200/libh5deltarice_b200.so; echo "== $v"; 
timeout 300 python tools/enc_time.py 153391 3500 4 2000 20 2>&1 | tail -1 | cut -c1-90
timeout 300 python tools/enc_time.py 76695 7000 8 2000 20 2>&1 | tail -1 | cut -c1-90
done
