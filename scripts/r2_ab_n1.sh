#!/bin/bash
bash scripts/r2_ab.sh "n1" u nofin f1 f3
echo "--- LDG tile kernel"
FLMIP_NO_TMA_TILES=1 bash scripts/r2_ab.sh "n1 n2" u
