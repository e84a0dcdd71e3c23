from .helpers import class_to_dict, get_load_path, get_args, set_seed, update_class_from_dict, export_policy_as_jit
from .logger import Logger
from .task_registry import task_registry
from .math import *
