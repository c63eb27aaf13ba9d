#!/bin/bash
# Launcher with the reference's parameter block (scripts/test_ttl.sh), for the B200 path.
#   bash scripts/test_ttl.sh A/R            # one GPU
#   NGPU=8 bash scripts/test_ttl.sh A/R     # one process per GPU, test samples sharded by index, one final all-reduce
# One deliberate difference from the reference's script: the data root is passed as the positional DIR
# argument; the reference's `--data $DATA_ROOT` is parsed by argparse as an abbreviation of --dataset_mode and the root
# silently stays at its default.
# DEYO_SELECTION defaults to True like the reference's script (weighted-entropy head, deyo.py); DEYO_SELECTION='' selects the
# confidence-selection + marginal-entropy head the paper describes.  Any non-empty string ("False" too) is truthy (Q1).
set -euo pipefail
cd "$(dirname "$0")/.."

DATA_ROOT=${DATA_ROOT:-/data/datasets}
TEST_SETS=${1:-A}            # A/V/R/K/I, joined with '/'
MODE='test'
ARCH=${ARCH:-ViT-B/16}       # ViT-B/16 | ViT-L/14
BS=64                        # views per test image
CTX_INIT='a_photo_of_a'
LR=5e-3
TTA_STEPS=${TTA_STEPS:-1}
PRINT_FRQ=200
GPU=${GPU:-0}
SELECTION_P=0.1              # 0.1 -> 6 of 64 views
LAYER_RANGE=${LAYER_RANGE:-9,11}
INIT_METHOD='xavier'
LORA_ENCODER='image'
RANK=16
DEYO_SELECTION=${DEYO_SELECTION-True}
NGPU=${NGPU:-1}

ARGS=("$DATA_ROOT" --test_sets "$TEST_SETS" --dataset_mode "$MODE" --arch "$ARCH" --b "$BS" --ctx_init "$CTX_INIT"
      --lr "$LR" --tta_steps "$TTA_STEPS" --print_freq "$PRINT_FRQ" --selection_p "$SELECTION_P"
      --layer_range "$LAYER_RANGE" --init_method "$INIT_METHOD" --lora_encoder "$LORA_ENCODER" --rank "$RANK"
      --deyo_selection "$DEYO_SELECTION" --views_on_device)

if [ "$NGPU" -gt 1 ]; then
  exec python3 -m torch.distributed.run --nnodes=1 --nproc-per-node "$NGPU" --master-addr 127.0.0.1 \
       --master-port "${MASTER_PORT:-29511}" ttl.py "${ARGS[@]}"
else
  exec python3 ttl.py "${ARGS[@]}" --gpu "$GPU"
fi
