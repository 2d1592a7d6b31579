#!/bin/bash
# Builds the prepared tuning variants HERE (no GPU needed) so that one gpurun call can A/B them:
#   scripts/round2_ab.sh                       # build into build_variants/ (travels with the snapshot; *.so is git-ignored)
#   gpurun --timeout 900 -- 'scripts/ab_variants.sh packed56 default build_variants/prefetch.so build_variants/inline_setup.so build_variants/inline_both.so; \
#                            scripts/ab_variants.sh ref96 default build_variants/prefetch.so build_variants/inline_setup.so build_variants/inline_both.so'
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
python -m mesh2splat_b200.build --out=build_variants/prefetch.so --def=M2S_FRAG_PREFETCH
python -m mesh2splat_b200.build --out=build_variants/inline_setup.so --def=M2S_INLINE_SETUP
python -m mesh2splat_b200.build --out=build_variants/inline_both.so --def=M2S_INLINE_SETUP --def=M2S_INLINE_RASTER
ls -la build_variants
