for lib in ph_steep ph_flat ph_flat2; do echo == $lib; MOOG_PROFILE_PHASES=1 MOOG_B200_LIB=$PWD/build/$lib.so python scripts/phase_profile.py 2>&1 | tail -6 | head -1; done
LIBS="lib_steep.so lib_flat2.so" bash scripts/run_ab_libs.sh
