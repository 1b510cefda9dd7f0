set -x
bash tools/gpu_final.sh
bash tools/gpu_profile.sh
