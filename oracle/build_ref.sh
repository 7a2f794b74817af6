#!/bin/bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference (haasn/libplacebo) from the sources
# where they lie under /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box).
#
# No reference source is copied into the repo: generated files (config.h, GLSL-preprocessed
# shader generators) and objects live only under oracle/_ref/.
#
# What is built: every C file of src/meson.build:91-114 (T0 math, gpu.c, dummy backend,
# dispatch, shader generators run through tools/glsl_preproc) -> oracle/_ref/libplacebo_ref.so
# The pixel arithmetic of the reference only exists as GLSL text and cannot execute here
# (no Vulkan/GL); the .so gives us (a) the real CPU parameter derivation (filter LUTs, tone
# curves, gamut 3DLUTs, matrices, grain tables, blue noise) and (b) the dummy GPU through which
# oracle/harvest.c pulls the exact LUT textures / uniforms / constants each pl_shader_* binds.
set -euo pipefail
R=${PL_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
W=$HERE/_ref
if [ ! -d "$R/src" ]; then
    echo "build_ref: $R not present; keeping prebuilt $W" >&2
    exit 0
fi
mkdir -p "$W/inc/libplacebo" "$W/gen/shaders" "$W/obj"
cd "$W"
sed -e 's/@majorver@/7/' -e 's/@apiver@/365/' -e 's/@extra_defs@/#define PL_HAVE_DOVI 1/' \
    "$R/src/include/libplacebo/config.h.in" > inc/libplacebo/config.h
printf '#define BUILD_API_VER 365\n#define BUILD_FIX_VER 0\n#define PL_HAVE_EXECINFO 1\n' > config_internal.h
echo '#define BUILD_VERSION "v7.365.0-ref"' > version.h
export PYTHONPATH=$R/3rdparty/jinja/src:$R/3rdparty/markupsafe/src:${PYTHONPATH:-}
PY=${PYTHON:-python3}
for f in renderer.c shaders.c; do
    [ gen/$f -nt "$R/src/$f" ] || $PY "$R/tools/glsl_preproc/main.py" "$R/src/$f" gen/$f
done
for f in colorspace custom custom_mpv deinterlacing dithering film_grain film_grain_av1 \
         film_grain_h274 icc lut sampling; do
    [ gen/shaders/$f.c -nt "$R/src/shaders/$f.c" ] || \
        $PY "$R/tools/glsl_preproc/main.py" "$R/src/shaders/$f.c" gen/shaders/$f.c
done
CF="-O2 -fPIC -D_GNU_SOURCE -DPL_HAVE_PTHREAD -DPTHREAD_HAS_SETCLOCK -DPL_STATIC -DPL_HAVE_DOVI \
    -fno-math-errno -I$W -I$W/inc -I$R/src/include -I$R/src -I$R/src/shaders \
    -I$R/3rdparty/fast_float/include -I$R/3rdparty/Vulkan-Headers/include -w"
SRCS="cache colorspace common dither dispatch dummy filters format gamut_mapping glsl/spirv gpu
      gpu/utils log options pl_alloc pl_string swapchain tone_mapping utils/dolbyvision
      utils/frame_queue utils/upload vulkan/stubs opengl/stubs"
OBJS=""
for s in $SRCS; do
    o=obj/$(echo $s | tr / _).o
    [ $o -nt "$R/src/$s.c" ] || gcc -std=c11 $CF -c "$R/src/$s.c" -o $o &
    OBJS="$OBJS $o"
done
for s in renderer shaders; do
    o=obj/gen_$s.o
    [ $o -nt gen/$s.c ] || gcc -std=c11 $CF -c gen/$s.c -o $o &
    OBJS="$OBJS $o"
done
for f in colorspace custom custom_mpv deinterlacing dithering film_grain film_grain_av1 \
         film_grain_h274 icc lut sampling; do
    o=obj/gen_shaders_$f.o
    [ $o -nt gen/shaders/$f.c ] || gcc -std=c11 $CF -c gen/shaders/$f.c -o $o &
    OBJS="$OBJS $o"
done
[ obj/convert.o -nt "$R/src/convert.cc" ] || g++ -std=c++20 $CF -c "$R/src/convert.cc" -o obj/convert.o &
wait
g++ -shared -Wl,-Bsymbolic -o libplacebo_ref.so $OBJS obj/convert.o -lm -lpthread
# record where the headers were so harvest.c can be compiled against them
echo "$CF" > cflags.txt
echo "build_ref: OK -> $W/libplacebo_ref.so"
# the harvester shim (our code, compiled against the reference headers in place)
gcc -std=c11 $CF -shared -Wl,-Bsymbolic -o "$W/libplref_harvest.so" "$HERE/harvest.c" -L"$W" -lplacebo_ref \
    -Wl,-rpath,'$ORIGIN' -lm -lpthread
echo "build_ref: OK -> $W/libplref_harvest.so"
