for L in fake_spectra_b200/libfsb200.so fake_spectra_b200/libfsb200_oldcolden.so; do
FSB200_LIB=$PWD/$L python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$L', json.dumps(d['colden']))"
done
