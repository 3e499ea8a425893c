python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for p in 0 1 0 1; do echo "== VASR_PDL=$p"; VASR_PDL=$p python tools/step_profile.py | tail -1; done
