timeout 400 python scripts/r2_mimi.py check time 2>&1 | grep -E "^tc|decode 2x|encode n|Error|error|rc="
