"""Cosine-weighted ISS (reference: ``fruits/iss/cos.py``).

Listed as the first "next" row of SURVEY.md section 8(f): it is used by
slices 2-3 of the reduced / general experiment fruits but is not part of the
north-star hot path.  Not built yet -- constructing it raises."""


class CosWISS:

    def __init__(self, *args, **kwargs) -> None:
        raise NotImplementedError(
            "CosWISS is not built yet (SURVEY.md section 8(f), rank 1)")
