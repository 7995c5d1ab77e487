from .BaseModule import BaseModule
