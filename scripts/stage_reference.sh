#!/bin/bash
# Stage the Python files of the reference's hot path next to the repo for ONE gpurun call, so that the GPU box can
# (a) run the unmodified train.py on the B200 implementation (tests/test_gpu_train_py.py) and (b) time the reference
# itself on the same GPU (scripts/reference_gpu_step.py).  baseline/_ref/ is git-ignored: nothing staged here is
# ever committed.  `scripts/stage_reference.sh clean` removes the copy again.
set -e
cd "$(dirname "$0")/.."
DST=baseline/_ref/IDEAS
if [ "$1" = "clean" ]; then rm -rf baseline/_ref; echo "removed baseline/_ref"; exit 0; fi
SRC=${IDEAS_REFERENCE:-/root/reference}
mkdir -p $DST/stylegan2/op
cp $SRC/train.py $SRC/models.py $SRC/utils.py $DST/
cp $SRC/stylegan2/model.py $DST/stylegan2/
cp $SRC/stylegan2/op/*.py $SRC/stylegan2/op/*.cpp $SRC/stylegan2/op/*.cu $DST/stylegan2/op/
echo "staged $(find $DST -type f | wc -l) files under $DST"
