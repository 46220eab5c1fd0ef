bash profiles/scripts/r02_final.sh 2>&1 | grep -v "^+"
bash profiles/scripts/r02_all.sh 2>&1 | grep -v "^+" | tail -14
