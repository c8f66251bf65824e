for wt in 8 16 32; do echo "WT=$wt"; B200_BAND_WT=$wt python profiles/scripts/band_probe.py --steps 10 2>&1 | grep '"device"' ; done
python -m pytest tests/test_gpu_band.py -x -q 2>&1 | tail -3
