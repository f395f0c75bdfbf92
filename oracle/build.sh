#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds the parity checker:
#   (1) oracle/_build/libjdsp_oracle.so  - the plain-C restatement (oracle/jdsp_oracle.c), always;
#   (2) oracle/_ref/*                    - the UNMODIFIED reference sources compiled where they lie
#       under $JDSP_REFERENCE_DIR (default /root/reference), only when that directory exists.
#       Nothing but binaries is written into the repo: preset variants are produced by piping
#       `sed` straight into g++ (stdin), never by saving an edited copy of a reference file.
# The product library (jeicyboodsp_b200/libjdsp.so) never links, loads or calls any of this.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${JDSP_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
BLD="$HERE/_build"
CXX="${CXX:-g++}"
CC="${CC:-gcc}"
OPT="-O3"
mkdir -p "$BLD"

echo "[oracle] building C restatement -> $BLD/libjdsp_oracle.so"
$CC $OPT -std=gnu11 -fPIC -shared -Wall -Wextra -o "$BLD/libjdsp_oracle.so" "$HERE/jdsp_oracle.c" -lm

if [ ! -d "$REF" ]; then
    echo "[oracle] $REF not present: keeping prebuilt oracle/_ref (if any)"
    exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
SHIM="$HERE/fftw_shim"
DRV="$HERE/drivers"
W="-w"   # the reference has 1-2 harmless warnings per file (void main, unused vars)

$CXX $OPT -c "$DRV/run_main.cpp" -o "$TMP/run_main.o"

# seds <file> <expr>...  : stream-edit a reference file to stdout and insist every expression hit
seds() {
    local f="$1"; shift
    local args=()
    for e in "$@"; do
        local pat="${e#s/}"; pat="${pat%%/*}"
        grep -Eq -- "$pat" "$f" || { echo "[oracle] sed pattern '$pat' not found in $f" >&2; exit 1; }
        args+=(-E -e "$e")
    done
    sed "${args[@]}" "$f"
}

# ---- FFTAlgorithm_ver2: round-trip programs and FFTProcess doorways -----------------------------
mkdir -p "$TMP/empty" && : > "$TMP/empty/fftw3.h"       # every FFTW line of this file is a comment
build_fft() {   # $1 = BLOCK_LEN
    local n="$1" src="$REF/FFTAlgorithm_ver2.cpp"
    if [ "$n" = 512 ]; then
        $CXX $OPT $W -fPIC -I"$TMP/empty" -Dmain=ref_main -c "$src" -o "$TMP/fft_$n.o"
    else
        seds "$src" "s/^#define BLOCK_LEN 512/#define BLOCK_LEN $n/" |
            $CXX $OPT $W -fPIC -I"$TMP/empty" -Dmain=ref_main -x c++ -c - -o "$TMP/fft_$n.o"
    fi
    if [ "$n" = 512 ] || [ "$n" = 1024 ]; then
        $CXX "$TMP/fft_$n.o" "$TMP/run_main.o" -o "$OUT/fft_roundtrip_$n" -lm
    fi
    objcopy --redefine-sym printf=jref_quiet_printf --redefine-sym puts=jref_quiet_puts --redefine-sym __printf_chk=jref_quiet_printf_chk "$TMP/fft_$n.o" "$TMP/fftq_$n.o"
    $CXX $OPT -fPIC -DJDSP_BLOCK_LEN=$n -c "$DRV/fftprocess_export.cpp" -o "$TMP/fftx_$n.o"
    $CXX -shared "$TMP/fftq_$n.o" "$TMP/fftx_$n.o" -o "$OUT/libfftprocess_$n.so" -lm
}
for n in 256 512 1024 2048 4096 8192 16384 32768; do build_fft $n; done

# ---- SpectralSubtraction_final / WienerFilter_final ----------------------------------------------
DEN_BENCH=(
    "s/^#define FFT_PROCESSING_SIZE 1024/#define FFT_PROCESSING_SIZE 512/"
    "s/^#define BLOCK_LEN 512/#define BLOCK_LEN 256/"
    "s/^#define KEEP_LEN 512/#define KEEP_LEN 256/"
    "s/0\.54 - 0\.46 \*/0.5 - 0.5 */g"
    "s/^#define THRESHOLD_OF_ZCR 200\.0/#define THRESHOLD_OF_ZCR 64.0/"
)
for prog in SpectralSubtraction_final WienerFilter_final; do
    short=ss; [ "$prog" = WienerFilter_final ] && short=wiener
    $CXX $OPT $W -I"$SHIM" -Dmain=ref_main -c "$REF/$prog.cpp" -o "$TMP/${short}_ref.o"
    $CXX "$TMP/${short}_ref.o" "$TMP/run_main.o" -o "$OUT/${short}_ref" -lm
    seds "$REF/$prog.cpp" "${DEN_BENCH[@]}" |
        $CXX $OPT $W -I"$SHIM" -Dmain=ref_main -x c++ -c - -o "$TMP/${short}_bench.o"
    $CXX "$TMP/${short}_bench.o" "$TMP/run_main.o" -o "$OUT/${short}_bench" -lm
done

# ---- Fast_Convolution_Based_3DAudio_Impl ---------------------------------------------------------
FC="$REF/Fast_Convolution_Based_3DAudio_Impl.cpp"
$CXX $OPT $W -I"$SHIM" -Dmain=ref_main -c "$FC" -o "$TMP/fc_ref.o"
$CXX $OPT -DJDSP_FILTER_LENGTH=7169 -c "$DRV/fastconv_taps_main.cpp" -o "$TMP/fc_ref_drv.o"
$CXX "$TMP/fc_ref.o" "$TMP/fc_ref_drv.o" -o "$OUT/fastconv_ref" -lm
# bench preset: block 512, N 1024, one history block, 512 taps (+1 zero); taps come from a file at run time
mkdir -p "$TMP/fc_bench"
printf '#define FILTER_LENGTH 513\n#define MAX_QUEUE_SIZE 1\ndouble rgdFirLPF_coefficients[FILTER_LENGTH] = {0};\n' \
    > "$TMP/fc_bench/FilterCoefficient.h"
( cd "$TMP/fc_bench" &&
  seds "$FC" "s/^#define BLOCK_SIZE 1024/#define BLOCK_SIZE 512/" \
             "s/^#define FFT_PROCESSING_SIZE 8192/#define FFT_PROCESSING_SIZE 1024/" |
      $CXX $OPT $W -I"$SHIM" -iquote "$TMP/fc_bench" -Dmain=ref_main -x c++ -c - -o "$TMP/fc_bench.o" )
$CXX $OPT -DJDSP_FILTER_LENGTH=513 -c "$DRV/fastconv_taps_main.cpp" -o "$TMP/fc_bench_drv.o"
$CXX "$TMP/fc_bench.o" "$TMP/fc_bench_drv.o" -o "$OUT/fastconv_bench" -lm

# ---- MFCCFeatureExtraction_auto_version1 ---------------------------------------------------------
MF="$REF/MFCCFeatureExtraction_auto_version1.cpp"
$CXX $OPT $W -I"$SHIM" -Dmain=ref_main -c "$MF" -o "$TMP/mfcc_ref.o"
$CXX "$TMP/mfcc_ref.o" "$TMP/run_main.o" -o "$OUT/mfcc_ref" -lm
# "mid" preset: 512/256 framing, 26 bands, 13 cepstra, 8 kHz mel span, through the reference's own code
seds "$MF" "s/^#define MFCC_LEN 12/#define MFCC_LEN 13/" \
           "s/^#define BLOCK_LEN 1024/#define BLOCK_LEN 512/" \
           "s/^#define WINDOW_LEN 1024/#define WINDOW_LEN 512/" \
           "s/^#define KEEP_LEN 512/#define KEEP_LEN 256/" \
           "s/^#define CHANNEL 38/#define CHANNEL 26/" \
           "s/^#define HALF_SAMPLING_RATE 22050\.0/#define HALF_SAMPLING_RATE 8000.0/" |
    $CXX $OPT $W -I"$SHIM" -Dmain=ref_main -x c++ -c - -o "$TMP/mfcc_mid.o"
$CXX "$TMP/mfcc_mid.o" "$TMP/run_main.o" -o "$OUT/mfcc_mid" -lm

# ---- PitchEstimation_method1: results are the per-block printf lines on stdout -------------
PT="$REF/PitchEstimation_method1.cpp"
$CXX $OPT $W -I"$SHIM" -Dmain=ref_main -c "$PT" -o "$TMP/pitch_ref.o"
$CXX "$TMP/pitch_ref.o" "$TMP/run_main.o" -o "$OUT/pitch_ref" -lm

# ---- BeamForming_MVDR_ver1: needs Eigen (absent) -> oracle/eigen_shim serves MatrixXcd ---------------
MV="$REF/BeamForming_MVDR_ver1.cpp"
$CXX $OPT $W -I"$SHIM" -I"$HERE/eigen_shim" -Dmain=ref_main -c "$MV" -o "$TMP/mvdr_ref.o"
$CXX "$TMP/mvdr_ref.o" "$TMP/run_main.o" -o "$OUT/mvdr_ref" -lm

cat > "$OUT/README.txt" <<EOF
Built by oracle/build.sh from the unmodified sources in $REF (g++ $($CXX -dumpversion), $OPT).
FFT behind the FFTW call sites: oracle/fftw_shim/fftw3.h (radix-2, double, exact pi) - NOT FFTW.
MatrixXcd behind BeamForming_MVDR_ver1: oracle/eigen_shim (Gauss-Jordan, partial pivoting) - NOT Eigen.
Binaries only; git-ignored; travels to the GPU box with gpurun.
EOF
echo "[oracle] reference binaries -> $OUT"
ls "$OUT"
