import json, sys
sys.path.insert(0, ".")
import torch
from bench import aux_env_step_api
print(json.dumps(aux_env_step_api(torch.device("cuda:0"))))
