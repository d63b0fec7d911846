"""Boundary types.  When the host application (the reference) is importable its own pydantic DTOs are
used, so objects cross the plug-in boundary unchanged; otherwise field-for-field mirrors of
app/core/dto.py:38-55 (RetrievalIn / Hit / RetrievalOut) keep the package usable stand-alone."""
from __future__ import annotations

from typing import Any, Dict, List

try:  # pragma: no cover - depends on the host application being on sys.path
    from app.core.dto import Hit, RetrievalIn, RetrievalOut  # type: ignore
    HOST_DTOS = True
except Exception:  # noqa: BLE001
    from pydantic import BaseModel, Field

    HOST_DTOS = False

    class RetrievalIn(BaseModel):
        query: str
        graph_id: str
        top_k: int = 20
        trace_id: str

    class Hit(BaseModel):
        id: str
        score: float
        meta: Dict[str, Any] = Field(default_factory=dict)

    class RetrievalOut(BaseModel):
        hits: List[Hit] = Field(default_factory=list)
        diagnostics: Dict[str, Any] = Field(default_factory=dict)
