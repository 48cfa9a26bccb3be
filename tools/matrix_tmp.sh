export MCSCE_B200_TILE_ROWS=2 MCSCE_B200_TILE_PACKED=1
python -m pytest tests -m gpu -q -x -k "fullsize or parity or scale or fuzz" 2>&1 | tail -3
python tools/strong_curve.py cfg3 2048,1024,512,256,128 2>&1 | grep "C="
python tools/strong_curve.py cfg4 512,64 2>&1 | grep "C="
python tools/strong_curve.py cfg5 32 2>&1 | grep "C="
