"""Mirror of the reference's model factory (src/models/ModelFactory.py:10-22): resolves
`configs['model']['name']` ("<ClassName>NN") to the class `<ClassName>` of the module of that name inside
this package and constructs it with `(configs, model_configs)`.  Unknown names raise RuntimeError, as upstream."""
import importlib
import inspect


def get_model(configs: dict, model_configs: dict = None):
    filename = configs['model']['name']
    classname = filename[:-2]
    try:
        module = importlib.import_module(f'{__package__}.{filename}')
    except ModuleNotFoundError as e:
        raise RuntimeError(f'Unknown model: {filename}') from e
    for name, cls in inspect.getmembers(module, inspect.isclass):
        if name == classname:
            return cls(configs, model_configs)
    raise RuntimeError(f'Unknown model: {filename}')
