"""Drop-in model classes; module and class names follow `get_model` (FoodRec/utils/utils.py:27-40):
`models/<name.lower()>.py` holds class `<name>`."""
