"""Argument validation with the semantics of the reference's ``type_checks.check_valid`` (type_checks.py:86-100):
returns the value cast to ``expected_type``; bounds are [lower, upper); a bound of 0/None is not enforced."""
import numbers

import numpy as np
import torch


class OutOfBoundsException(Exception):
    def __init__(self, value, lower_bound, upper_bound):
        super().__init__(f"Value {value} was not in bounds: [{lower_bound}, {upper_bound}).")


class InvalidTypeException(Exception):
    def __init__(self, value, expected_type):
        super().__init__(f"Value {value} was of type {type(value)} but expected to be of type {expected_type} "
                         f"(or a subclass of this type) .")


def _is_kind(value, expected_type) -> bool:
    if isinstance(value, torch.Tensor):
        return value.dtype.is_floating_point == (expected_type is float)
    if isinstance(value, np.generic):
        return isinstance(value, np.floating if expected_type is float else np.integer)
    if isinstance(value, bool):
        return False
    if expected_type is float:
        return isinstance(value, float)
    return isinstance(value, numbers.Integral)


def check_valid(value, expected_type, lower_bound=None, upper_bound=None, allow_none: bool = False):
    if value is None:
        if allow_none:
            return None
        raise Exception(f"Invalid input: Got None, but expected type {expected_type}.")
    if expected_type not in (int, float):
        raise Exception(f"Unexpected data type, must be either int or float, but was {expected_type}")
    if not isinstance(value, (torch.Tensor, np.generic, int, float)):
        raise Exception(f"Unsupported type ({type(value)}) for typecheck.")
    if not _is_kind(value, expected_type):
        raise InvalidTypeException(value, expected_type)
    # the reference only enforces truthy bounds (type_checks.py:27-35)
    if lower_bound and not value >= expected_type(lower_bound):
        raise OutOfBoundsException(value, lower_bound, upper_bound)
    if upper_bound and not value < expected_type(upper_bound):
        raise OutOfBoundsException(value, lower_bound, upper_bound)
    return expected_type(value)
