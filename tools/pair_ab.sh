for c in C1_perm_grid3 C2_lf8_line C4_pauli10_line C3_clifford8_full; do
for pr in 1 0; do
  QG_ENGINE_LIB=$PWD/qiskit_gym_b200/libqg_engine_knobs.so QG_PAIR=$pr timeout 200 python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline --no-synth --no-collector --no-e2e --no-per-step --no-packed 2>/dev/null | python -c "
import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c pair=$pr value %.3e frac %.3f'%(b['value'], b['roofline']['frac']))"
done; done
