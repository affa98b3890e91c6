set -x
bash scripts/gpu_final.sh
bash scripts/r2_profile.sh
