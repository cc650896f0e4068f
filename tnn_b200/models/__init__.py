from .zoo import ExampleModels, load_or_create_model  # noqa: F401
