"""Conversation templates of the TRACE surface (mirrors trace/conversation.py: Conversation.get_prompt LLAMA_2
branch :78-98, conv_llama_2 :383-394, conv_templates :501-520).  Pure host string work; only the templates the
inference drivers use are provided ('llama_2' is the one every TRACE driver selects; 'v1'/'plain' kept for the
drivers' `--conv-mode` flag).  Prompt text is observable behaviour (it becomes token ids), so the system prompts
are the same public LLaVA / Llama-2 strings."""
from __future__ import annotations

import dataclasses
from enum import Enum, auto
from typing import List


class SeparatorStyle(Enum):
    SINGLE = auto()
    TWO = auto()
    MPT = auto()
    PLAIN = auto()
    LLAMA_2 = auto()
    QWEN = auto()


@dataclasses.dataclass
class Conversation:
    system: str
    roles: tuple
    messages: List[List[str]]
    offset: int
    sep_style: SeparatorStyle = SeparatorStyle.SINGLE
    sep: str = "###"
    sep2: str = None
    version: str = "Unknown"
    skip_next: bool = False

    def get_prompt(self) -> str:
        msgs = self.messages
        if self.sep_style == SeparatorStyle.LLAMA_2:
            out = ""
            for i, (role, message) in enumerate(msgs):
                if i == 0:
                    assert message, "first message should not be none"
                    assert role == self.roles[0], "first message should come from user"
                if not message:
                    continue
                if isinstance(message, tuple):
                    message = message[0]
                if i == 0:
                    message = f"<<SYS>>\n{self.system}\n<</SYS>>\n\n" + message
                if i % 2 == 0:
                    out += self.sep + f"[INST] {message} [/INST]"
                else:
                    out += " " + message + " " + self.sep2
            return out.lstrip(self.sep)
        if self.sep_style == SeparatorStyle.TWO:
            seps = [self.sep, self.sep2]
            out = self.system + seps[0]
            for i, (role, message) in enumerate(msgs):
                if message:
                    if isinstance(message, tuple):
                        message = message[0]
                    out += role + ": " + message + seps[i % 2]
                else:
                    out += role + ":"
            return out
        if self.sep_style == SeparatorStyle.SINGLE:
            out = self.system + self.sep
            for role, message in msgs:
                if message:
                    if isinstance(message, tuple):
                        message = message[0]
                    out += role + ": " + message + self.sep
                else:
                    out += role + ":"
            return out
        if self.sep_style == SeparatorStyle.PLAIN:
            seps = [self.sep, self.sep2]
            out = self.system
            for i, (role, message) in enumerate(msgs):
                if message:
                    if isinstance(message, tuple):
                        message = message[0]
                    out += message + seps[i % 2]
            return out
        raise ValueError(f"Invalid style: {self.sep_style}")

    def append_message(self, role, message):
        self.messages.append([role, message])

    def copy(self):
        return Conversation(system=self.system, roles=self.roles, messages=[[x, y] for x, y in self.messages],
                            offset=self.offset, sep_style=self.sep_style, sep=self.sep, sep2=self.sep2, version=self.version)


conv_llama_2 = Conversation(
    system="You are a helpful, respectful and honest assistant. Always answer as helpfully as possible, while being safe.  "
           "Your answers should not include any harmful, unethical, racist, sexist, toxic, dangerous, or illegal content. "
           "Please ensure that your responses are socially unbiased and positive in nature.\n\n"
           "If a question does not make any sense, or is not factually coherent, explain why instead of answering something "
           "not correct. If you don't know the answer to a question, please don't share false information.",
    roles=("USER", "ASSISTANT"), version="llama_v2", messages=[], offset=0,
    sep_style=SeparatorStyle.LLAMA_2, sep="<s>", sep2="</s>")

conv_vicuna_v1 = Conversation(
    system="A chat between a curious user and an artificial intelligence assistant. "
           "The assistant gives helpful, detailed, and polite answers to the user's questions.",
    roles=("USER", "ASSISTANT"), version="v1", messages=[], offset=0,
    sep_style=SeparatorStyle.TWO, sep=" ", sep2="</s>")

conv_plain = Conversation(system="", roles=("", ""), messages=[], offset=0, sep_style=SeparatorStyle.PLAIN, sep="", sep2="\n")

default_conversation = conv_vicuna_v1
conv_templates = {"default": conv_vicuna_v1, "v1": conv_vicuna_v1, "vicuna_v1": conv_vicuna_v1,
                  "llama_2": conv_llama_2, "plain": conv_plain}
