"""Data-preprocessor plugin factory with the reference's contract
(src/data_preprocessors/DataPreprocessorFactory.py:13-26): `configs['data_loader']['data_preprocessor_name']` names a
module `<Name>NN` of this package whose class `<Name>` is constructed as `<Name>(configs, mode, raw_data_dict,
model_configs)`."""
import importlib
import inspect
from typing import Optional


def get_data_preprocessor(configs: dict, mode: str, *, raw_data_dict: Optional[dict] = None,
                          model_configs: Optional[dict] = None):
    filename = configs['data_loader']['data_preprocessor_name']
    classname = filename[:-2]
    try:
        module = importlib.import_module(f'{__package__}.{filename}')
    except ModuleNotFoundError as e:
        raise RuntimeError(f'Unknown data preprocessor: {filename}') from e
    for name, cls in inspect.getmembers(module, inspect.isclass):
        if name == classname:
            return cls(configs, mode, raw_data_dict, model_configs)
    raise RuntimeError(f'Unknown data preprocessor: {filename}')
