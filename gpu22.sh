cp nfisam_b200/libnfisam_b200.so scratch/lib_orig.so
for v in orig mb8 mb10 mb12; do
cp scratch/lib_$v.so nfisam_b200/libnfisam_b200.so
python bench.py --steps 20 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$v', round(j['ms_per_step'],3), round(j['roofline']['frac'],3))"
done
cp scratch/lib_orig.so nfisam_b200/libnfisam_b200.so
