# Run HERE (build container) at the start of a round: recreate the worktrees of the unverified branches under .wip/
# (git-ignored; they travel to the GPU box with the snapshot) and build their libraries.  Then:
#   gpurun --timeout 900 -- 'bash scripts/gpu_wip.sh'
set -e
cd "$(dirname "$0")/.."
git worktree prune
for b in attn swinir; do
  if [ ! -d .wip/$b ]; then git worktree add .wip/$b $b-wip; fi
  (cd .wip/$b && git merge -q main -m "merge main" || echo "!! merge conflict in $b-wip: resolve by hand" && python -c "
import sys; sys.path.insert(0, '.')
from edtr_b200 import lib
lib.build(force=True)
print('built', lib.LIB_PATH)")
done
