run() { echo "== $*"; env "$@" timeout 250 python tools/prof_v4.py $S 2>&1 | grep -v "^{" | tail -1 ; }
S=256 run DG_V4_SLOG=10 DG_V4_NCW=8 DG_V4_SLOT=4096
S=256 run DG_V4_SLOG=10 DG_V4_NCW=10 DG_V4_SLOT=4096
S=256 run DG_V4_SLOG=10 DG_V4_NCW=6 DG_V4_SLOT=2048 DG_V4_NSLOT=6
