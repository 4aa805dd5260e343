// stub: header removed from recent PyTorch; nothing from it is used
