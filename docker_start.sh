#!/usr/bin/env bash
# Start (or stop) the docker-compose deployment: two pipeline-stage workers and a coordinator (reference: docker_start.sh).
#   ./docker_start.sh              build the image if needed, start detached, follow the coordinator's log
#   ./docker_start.sh --clean      stop and remove the containers and the network
#   ./docker_start.sh --no-follow  start detached and return
set -euo pipefail
cd "$(dirname "$0")"
follow=1
compose() { if docker compose version > /dev/null 2>&1; then docker compose "$@"; else docker-compose "$@"; fi; }
while [[ $# -gt 0 ]]; do
  case "$1" in
    --clean | -c) compose down --remove-orphans; exit 0 ;;
    --no-follow) follow=0 ;;
    --help | -h) sed -n '2,5p' "$0" | sed 's/^# \{0,1\}//'; exit 0 ;;
    *) echo "unknown option: $1 (see --help)" >&2; exit 1 ;;
  esac
  shift
done
command -v docker > /dev/null || { echo "docker is not installed" >&2; exit 1; }
compose build
compose up -d
[[ $follow == 1 ]] && compose logs -f coordinator
